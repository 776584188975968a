set -x
mkdir -p gpurun_out
T=${1:-r02m}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
for f in bench_n2 bench bench_ref; do echo "== $f"; tail -n 3 gpurun_out/${T}_$f.err; cut -c1-300 gpurun_out/${T}_$f.json; done
