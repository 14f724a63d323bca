mkdir -p gpurun_out
export ABR_LIB_PATH=$PWD/aboria_b200/libC/libabr.so
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2m_bench_C.json 2> gpurun_out/r2m_bench_C.err
unset ABR_LIB_PATH
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_A.json 2> gpurun_out/r2m_bench_A.err
python - <<PY
import json
for v in ("C","A"):
    try:
        d=json.loads(open(f"gpurun_out/r2m_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, round(d["ms_per_step"],3), round(d["ms_build"],3), round(d["ms_matvec"],3), f'{d["value"]:.4g}', round(d["e2e"]["ms_per_step"],2))
        if d.get("extra"):
            for c in d["extra"]["configs_c1_c4"]: print("   ", c["config"][:40], round(c["ms_build"],3), round(c["ms_matvec"],3), f'{c["pairs_per_s"]:.3g}', c["rows_recomputed_by_exact_walk"])
    except Exception as e:
        print(v, "fail", e, open(f"gpurun_out/r2m_bench_{v}.err").read()[-800:])
PY
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2m_tests.log; tail -6 gpurun_out/r2m_tests.log
