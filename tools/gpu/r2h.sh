mkdir -p gpurun_out
NP=${NP:-2}
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP "$@" > gpurun_out/r2h_${name}_n$NP.json 2> gpurun_out/r2h_${name}_n$NP.err
  echo "$name rc=$?"; grep -E "Error|error|Traceback|assert" gpurun_out/r2h_${name}_n$NP.err | tail -5
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2h_${name}_n$NP.json").read().strip().splitlines()[-1])
    print("$name N=$NP", round(d["ms_per_step"],3), f'{d["value"]:.4g}', "e2e", round(d["e2e"]["ms_per_step"],2), d["config"]["n_particles_per_gpu"], d["per_rank"], d["config"]["y_checksum"])
except Exception as e:
    print("parse fail", e)
PY
}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/r2h_slab_check_n$NP.log 2>&1
echo "slab check rc=$?"; grep -cE "OK$" gpurun_out/r2h_slab_check_n$NP.log; grep -E "FAIL|Error|Traceback" gpurun_out/r2h_slab_check_n$NP.log | tail
run weak --steps 10 --warmup 3
run clustered --cloud clustered --steps 5 --warmup 3
run strong64 --scaling strong --n-total 64000000 --steps 10 --warmup 3
