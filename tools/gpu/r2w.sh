#!/bin/bash
# A/B of the drain's partial-sum update (ABR_ACC_MODE 0/1/2) on the c5 bench line, same box
cd "$GRAFT_REPO_ROOT" || exit 1
out=gpurun_out/r2w_acc_modes.txt
for rep in 1 2; do
for L in lib libM1 libM2; do
  r=$(ABR_LIB_PATH=$PWD/aboria_b200/$L/libabr.so timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['ms_build'], d['ms_matvec'], d['config']['pairs_per_matvec'])")
  echo "$L rep$rep: step/build/product ms, pairs: $r" | tee -a $out
done
done
