#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
NP=2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/r2y_slab_check_n$NP.log 2>&1
echo "slab check rc=$? ok-lines=$(grep -cE 'OK$' gpurun_out/r2y_slab_check_n$NP.log) fail-lines=$(grep -cE 'FAIL' gpurun_out/r2y_slab_check_n$NP.log)"; grep -E "FAIL|Error|Traceback" gpurun_out/r2y_slab_check_n$NP.log | tail -5
ABR_NCCL_TIMEOUT_S=60 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP --steps 8 --warmup 3 > gpurun_out/r2y_bench_weak_n$NP.json 2> gpurun_out/r2y_bench_weak_n$NP.err
echo "weak rc=$?"; grep -E "Error|error|assert" gpurun_out/r2y_bench_weak_n$NP.err | head -5
python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bench_weak_n2.json").read().strip().splitlines()[-1])
print("weak N=2 ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.4g}', "e2e ms", round(d["e2e"]["ms_per_step"],2))
PY
