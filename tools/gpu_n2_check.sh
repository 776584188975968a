# two-GPU check of the default sharded paths: gather tests, C5 (weak) and the default bench line (C2 strong scaling + c5 secondary)
set -x
mkdir -p gpurun_out
T=${1:-n2}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${T}_tests.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
tail -n 2 gpurun_out/${T}_bench_n2.err | cut -c1-200; cut -c1-200 gpurun_out/${T}_bench_n2.json
