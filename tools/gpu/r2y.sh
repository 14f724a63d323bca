#!/bin/bash
# staged record move of the two-level build (k_radix_scatter<2>): parity, then A/B on the c5 bench line
cd "$GRAFT_REPO_ROOT" || exit 1
out=gpurun_out/r2y_stage_records.txt
: > $out
timeout 600 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 | tee -a $out
for rep in 1 2; do
for S in 1 0; do
  r=$(ABR_STAGE_RECORDS=$S timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_build'], d['ms_matvec'], d['config']['pairs_per_matvec'])")
  echo "ABR_STAGE_RECORDS=$S rep$rep: step/build/product ms, pairs: $r" | tee -a $out
done
done
