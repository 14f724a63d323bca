mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"tiled_kernel|walk_kernel|k_pack" -s 6 -c 12 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2j_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); idi=hdr.index('ID')
by={}
for r in rows[1:]:
    by.setdefault(int(r[idi]),{'k':r[ki]})[r[mi]]=r[vi]
for i in sorted(by):
    d=by[i]; print(i, d['k'][:60], d.get('gpu__time_duration.sum'), d.get('smsp__inst_executed.sum'))
PY
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2j_tests.log; tail -8 gpurun_out/r2j_tests.log
