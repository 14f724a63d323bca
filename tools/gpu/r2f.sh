mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build or clustered or c5 or c2" 2>&1 | tail -25) > gpurun_out/r2f_tests.log
tail -25 gpurun_out/r2f_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_cs_ -s 8 -c 8 --csv --log-file gpurun_out/r2f_launches_build.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<PY
import json,csv
try:
    d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["ms_build"], d["ms_matvec"], d["value"])
except Exception as e:
    print("fail", e, open("gpurun_out/r2f_bench.err").read()[-1500:])
rows=[r for r in csv.reader(open('gpurun_out/r2f_launches_build.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
by={}
for r in rows[1:]:
    by.setdefault(int(r[idi]),{'k':r[ki]})[r[mi]]=r[vi]
for i in sorted(by):
    d=by[i]; print(i, d['k'][:40], {k:v for k,v in d.items() if k!='k'})
PY
