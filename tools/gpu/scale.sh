# usage: NP=<gpus> bash tools/gpu/scale.sh   — slab parity check + weak / strong (256M) / clustered (c4) bench lines at NP GPUs
mkdir -p gpurun_out
NP=${NP:-2}
TAG=${TAG:-r2s}
run() { # name, args...
  name=$1; shift
  if [ "$NP" = "1" ]; then
    timeout 1200 python bench.py --gpus 1 "$@" --no-cpu-baseline --no-extra > gpurun_out/${TAG}_${name}_n$NP.json 2> gpurun_out/${TAG}_${name}_n$NP.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $NP "$@" > gpurun_out/${TAG}_${name}_n$NP.json 2> gpurun_out/${TAG}_${name}_n$NP.err
  fi
  echo "$name rc=$?"; grep -E "Error|error|Traceback|assert" gpurun_out/${TAG}_${name}_n$NP.err | tail -5
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${name}_n$NP.json").read().strip().splitlines()[-1])
    print("$name N=$NP ms/step", round(d["ms_per_step"],3), "pairs/s", f'{d["value"]:.4g}', "e2e ms", round(d["e2e"]["ms_per_step"],2), "per_rank", d.get("per_rank"), "chk", d["config"].get("y_checksum"), "n", d["config"]["n_particles_per_gpu"])
except Exception as e:
    print("parse fail", e)
PY
}
if [ "$NP" != "1" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$NP --master-addr 127.0.0.1 --master-port 29541 tests/slab_gpu_check.py > gpurun_out/${TAG}_slab_check_n$NP.log 2>&1
  echo "slab check rc=$? ok-lines=$(grep -cE 'OK$' gpurun_out/${TAG}_slab_check_n$NP.log) fail-lines=$(grep -cE 'FAIL' gpurun_out/${TAG}_slab_check_n$NP.log)"; grep -E "FAIL|Error|Traceback" gpurun_out/${TAG}_slab_check_n$NP.log | tail
fi
run weak --steps 10 --warmup 3
run strong --scaling strong --steps 5 --warmup 3
run clustered --cloud clustered --steps 5 --warmup 3
