set -x
mkdir -p gpurun_out
T=${1:-r02h}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
for w in mlp128 c3 hmc_mlp128 c4; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_$w.json 2> gpurun_out/${T}_$w.err
tail -2 gpurun_out/${T}_$w.err; cut -c1-200 gpurun_out/${T}_$w.json
done
