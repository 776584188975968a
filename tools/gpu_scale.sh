set -x
mkdir -p gpurun_out
T=${1:-r02n}
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench_n$N.json 2> gpurun_out/${T}_bench_n$N.err
tail -n 2 gpurun_out/${T}_bench_n$N.err; cut -c1-200 gpurun_out/${T}_bench_n$N.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --c5-gather sm > gpurun_out/${T}_c5_n8_sm.json 2> gpurun_out/${T}_c5_n8_sm.err
cut -c1-200 gpurun_out/${T}_c5_n8_sm.json
