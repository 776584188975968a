set -x
mkdir -p gpurun_out
T=${1:-r02s}
N=${2:-8}
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pcd.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5_n${N}_skew.json 2> gpurun_out/${T}_c5_n${N}_skew.err
tail -n 2 gpurun_out/${T}_c5_n${N}_skew.err; cut -c1-200 gpurun_out/${T}_c5_n${N}_skew.json
EBM_B200_NO_SKEW=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5_n${N}_noskew.json 2> gpurun_out/${T}_c5_n${N}_noskew.err
tail -n 2 gpurun_out/${T}_c5_n${N}_noskew.err; cut -c1-200 gpurun_out/${T}_c5_n${N}_noskew.json
