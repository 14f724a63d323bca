mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2zz_tests.log; tail -3 gpurun_out/r2zz_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2zz_bench_n1.json 2> gpurun_out/r2zz_bench_n1.err; echo "bench rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r2zz_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2zz_ncu1.log 2>&1
timeout 300 python bench.py --scaling strong --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/r2zz_bench_strong_n1.json 2> gpurun_out/r2zz_bench_strong_n1.err; echo "strong rc=$?"
timeout 300 python bench.py --cloud clustered --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/r2zz_bench_clustered_n1.json 2> gpurun_out/r2zz_bench_clustered_n1.err; echo "clustered rc=$?"
python - <<PY
import json
for f in ("r2zz_bench_n1","r2zz_bench_strong_n1","r2zz_bench_clustered_n1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), f'{d["value"]:.4g}', d.get("ms_build"), d.get("ms_matvec"), d["e2e"].get("ms_per_step"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "fail", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
