mkdir -p gpurun_out
NP=2
ABR_SLAB_TEST_CAP=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/r2p_slab_check_cap_n$NP.log 2>&1
echo "slab check (rank-dependent reserve) rc=$? ok-lines=$(grep -cE 'OK$' gpurun_out/r2p_slab_check_cap_n$NP.log) fail-lines=$(grep -cE 'FAIL' gpurun_out/r2p_slab_check_cap_n$NP.log)"; grep -E "FAIL|Error|Traceback" gpurun_out/r2p_slab_check_cap_n$NP.log | tail -5
ABR_NCCL_TIMEOUT_S=60 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP --cloud clustered --steps 4 --warmup 3 > gpurun_out/r2p_clustered_n$NP.json 2> gpurun_out/r2p_clustered_n$NP.err
echo "clustered rc=$?"; grep -E "Error|error|assert" gpurun_out/r2p_clustered_n$NP.err | head -5
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2p_clustered_n2.json").read().strip().splitlines()[-1])
    print("clustered N=2 ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.4g}', "e2e ms", round(d["e2e"]["ms_per_step"],2), d["per_rank"], d["config"]["n_particles_per_gpu"], d["config"]["layers"])
except Exception as e:
    print("parse fail", e)
PY
