# bounded 8-GPU validation: slab parity check + weak / strong (256M) / clustered (c4) bench lines
mkdir -p gpurun_out
NP=8
TAG=r2s
export ABR_NCCL_TIMEOUT_S=60
run() { # name, timeout, args...
  name=$1; to=$2; shift 2
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP "$@" > gpurun_out/${TAG}_${name}_n$NP.json 2> gpurun_out/${TAG}_${name}_n$NP.err
  echo "$name rc=$?"; grep -E "Error|error|assert" gpurun_out/${TAG}_${name}_n$NP.err | head -3
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}_n$NP.json").read().strip().splitlines()[-1])
    print("$name N=$NP ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.4g}', "e2e ms", round(d["e2e"]["ms_per_step"],2), "build", [round(v,2) for v in d["per_rank"]["ms_build"]], "mv", [round(v,2) for v in d["per_rank"]["ms_matvec"]], "chk", d["config"].get("y_checksum"))
except Exception as e:
    print("parse fail", e)
PY
}
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/${TAG}_slab_check_n$NP.log 2>&1
echo "slab check rc=$? ok-lines=$(grep -cE 'OK$' gpurun_out/${TAG}_slab_check_n$NP.log) fail-lines=$(grep -cE 'FAIL' gpurun_out/${TAG}_slab_check_n$NP.log)"; grep -E "FAIL|Error|Traceback" gpurun_out/${TAG}_slab_check_n$NP.log | tail -3
run weak 85 --steps 10 --warmup 3
run strong 85 --scaling strong --steps 10 --warmup 3
run clustered 85 --cloud clustered --steps 5 --warmup 3
