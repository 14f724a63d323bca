mkdir -p gpurun_out
NP=${NP:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/r2g_slab_check_n$NP.log 2>&1
echo "slab check rc=$?"; grep -E "OK|FAIL|Error|error|Traceback" gpurun_out/r2g_slab_check_n$NP.log | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $NP --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$NP.json 2> gpurun_out/r2g_bench_n$NP.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2g_bench_n$NP.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2g_bench_n$NP.json").read().strip().splitlines()[-1])
    print("N=$NP", d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("bench parse fail", e)
PY
