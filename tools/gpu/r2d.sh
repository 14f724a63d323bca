mkdir -p gpurun_out
(timeout 700 python -m pytest tests/test_gpu_symmetric.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2d_tests.log
for v in 1 0; do ABR_SYMMETRIC=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench_s$v.json 2> gpurun_out/r2d_bench_s$v.err; done
tail -15 gpurun_out/r2d_tests.log
python - <<PY
import json
for v in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/r2d_bench_s{v}.json").read().strip().splitlines()[-1])
        print(v, d["ms_per_step"], d["ms_build"], d["ms_matvec"], d["value"])
    except Exception as e:
        print(v, "fail", e, open(f"gpurun_out/r2d_bench_s{v}.err").read()[-1500:])
PY
