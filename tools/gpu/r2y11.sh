#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
out=gpurun_out/r2y_stage_records.txt
echo "== enforce_one: inside-domain fast path" | tee -a $out
timeout 700 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_golden_fixtures.py -m gpu -q -x 2>&1 | tail -3 | tee -a $out
r=$(timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_build'], d['ms_matvec'], d['config']['pairs_per_matvec'])")
echo "default: step/build/product ms, pairs: $r" | tee -a $out
