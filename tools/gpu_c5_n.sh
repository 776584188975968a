# C5 on N GPUs: gather variants of the wide kernel's in-kernel pusher (multicast stores / bulk copies per peer / 16-byte stores per peer)
set -x
N=${1:-2}
T=${2:-c5}
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c5 --steps 10 --warmup 3 > gpurun_out/${T}_n${N}_${name}.json 2> gpurun_out/${T}_n${N}_${name}.err
  tail -n 2 gpurun_out/${T}_n${N}_${name}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_n${N}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", "ms_per_step", d["ms_per_step"], "gather_check", d.get("gather_check"), "value", d["value"])
except Exception as e:
    print("${name} FAILED", e)
PY
}
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5; fi
run bulk EBM_B200_NO_MULTICAST=1
run mc EBM_B200_PUSH_BULK=1
if [ "$N" = "2" ]; then run p2p EBM_B200_NO_MULTICAST=1 EBM_B200_PUSH_BULK=0; fi
