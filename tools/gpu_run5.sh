set -x
mkdir -p gpurun_out
T=r02d
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/${T}_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5_n2.json 2> gpurun_out/${T}_c5_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --c5-gather dma > gpurun_out/${T}_c5_n2_dma.json 2> gpurun_out/${T}_c5_n2_dma.err
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_c5_n1.json 2> gpurun_out/${T}_c5_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_tests.txt
for f in c5_n2 c5_n2_dma c5_n1 bench_n2 bench; do echo "== $f"; tail -3 gpurun_out/${T}_$f.err; cut -c1-600 gpurun_out/${T}_$f.json; done
