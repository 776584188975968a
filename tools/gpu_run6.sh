set -x
mkdir -p gpurun_out
T=${1:-r02e}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
timeout 300 python bench.py --workload mlp128 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_mlp128.json 2> gpurun_out/${T}_mlp128.err
EBM_B200_TC_SINGLE=1 timeout 300 python bench.py --workload mlp128 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_mlp128_single.json 2> gpurun_out/${T}_mlp128_single.err
timeout 300 python bench.py --workload mlp128_bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_mlp128_bf16.json 2> gpurun_out/${T}_mlp128_bf16.err
for f in mlp128 mlp128_single mlp128_bf16; do tail -3 gpurun_out/${T}_$f.err; cut -c1-330 gpurun_out/${T}_$f.json; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:langevin_mlp_tc2 -c 1 -o gpurun_out/${T}_mlp128_tc2 python bench.py --workload mlp128 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${T}_ncu.err
tail -3 gpurun_out/${T}_ncu.err
