#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
for A in 1 0; do
ABR_RECORD_AOS=$A timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"k_radix_scatter|k_gather|k_enforce|k_bound" -s 24 -c 8 --csv --log-file gpurun_out/r2y5_aos$A.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
done
ls -la gpurun_out/r2y5_*
