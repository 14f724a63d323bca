mkdir -p gpurun_out
for rep in 1 2 3; do
for v in C A; do
  if [ $v = C ]; then export ABR_LIB_PATH=$PWD/aboria_b200/libC/libabr.so; else unset ABR_LIB_PATH; fi
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2l_bench_$v$rep.json 2> gpurun_out/r2l_bench_$v$rep.err
done; done
unset ABR_LIB_PATH
python - <<PY
import json
for v in ("C1","A1","C2","A2","C3","A3"):
    try:
        d=json.loads(open(f"gpurun_out/r2l_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["ms_per_step"],3), round(d["ms_build"],3), round(d["ms_matvec"],3), f'{d["value"]:.4g}', round(d["e2e"]["ms_per_step"],2))
    except Exception as e:
        print(v, "fail", e, open(f"gpurun_out/r2l_bench_{v}.err").read()[-800:])
PY
