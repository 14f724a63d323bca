#!/bin/bash
# staged record move: timing with the hoisted tile-offset load, then one ncu --set full capture of the record movers
cd "$GRAFT_REPO_ROOT" || exit 1
out=gpurun_out/r2y_stage_records.txt
for S in 1 0; do
  r=$(ABR_STAGE_RECORDS=$S timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_build'], d['ms_matvec'], d['config']['pairs_per_matvec'])")
  echo "hoisted offsets, ABR_STAGE_RECORDS=$S: step/build/product ms, pairs: $r" | tee -a $out
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_radix_scatter|k_gather_fused" -s 8 -c 4 -f -o gpurun_out/r2y_build_movers python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2y_ncu.log 2>&1
tail -3 gpurun_out/r2y_ncu.log; ls -la gpurun_out/r2y_build_movers.ncu-rep
