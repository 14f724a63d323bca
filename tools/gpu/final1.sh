mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2z_tests.log; tail -4 gpurun_out/r2z_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; echo "bench rc=$?"
ABR_REF_BUDGET_S=40 timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_bench_reference_arm.err; echo "ref rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2z_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tiled_kernel -s 4 -c 1 -o gpurun_out/r2z_tiled python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2z_ncu2.log 2>&1
timeout 400 python bench.py --scaling strong --steps 3 --warmup 2 --no-cpu-baseline --no-extra > gpurun_out/r2z_bench_strong_n1.json 2> gpurun_out/r2z_bench_strong_n1.err; echo "strong rc=$?"
python - <<PY
import json
for f in ("r2z_bench_n1","r2z_bench_reference_arm","r2z_bench_strong_n1"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), f'{d["value"]:.4g}', d.get("ms_build"), d.get("ms_matvec"), d["e2e"].get("ms_per_step"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "fail", e, open(f"gpurun_out/{f}.err").read()[-600:])
PY
ls -la gpurun_out/r2z_tiled.ncu-rep
