#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
NP=4
ABR_NCCL_TIMEOUT_S=60 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP --steps 8 --warmup 3 > gpurun_out/r2zz_bench_weak_n$NP.json 2> gpurun_out/r2zz_bench_weak_n$NP.err
echo "weak rc=$?"; grep -E "Error|error|assert" gpurun_out/r2zz_bench_weak_n$NP.err | head -5
python - <<PY
import json
d=json.loads(open("gpurun_out/r2zz_bench_weak_n4.json").read().strip().splitlines()[-1])
print("weak N=4 ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.4g}', "e2e ms", round(d["e2e"]["ms_per_step"],2))
PY
