mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2i_tests.log
tail -12 gpurun_out/r2i_tests.log
./tools/build_floor > gpurun_out/r2i_build_floor.txt 2>&1; cat gpurun_out/r2i_build_floor.txt
for v in A B; do
  if [ $v = B ]; then export ABR_LIB_PATH=$PWD/aboria_b200/libB/libabr.so; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2i_bench_$v.json 2> gpurun_out/r2i_bench_$v.err
done
unset ABR_LIB_PATH
python - <<PY
import json
for v in "AB":
    try:
        d=json.loads(open(f"gpurun_out/r2i_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, d["ms_per_step"], d["ms_build"], d["ms_matvec"], d["value"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(v, "fail", e, open(f"gpurun_out/r2i_bench_{v}.err").read()[-1500:])
PY
