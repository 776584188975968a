set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_tests.txt
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r02a_c2.json 2> gpurun_out/r02a_c2.err
python bench.py --workload mlp128 --no-cpu-baseline > gpurun_out/r02a_mlp128.json 2> gpurun_out/r02a_mlp128.err
python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r02a_c3.json 2> gpurun_out/r02a_c3.err
python bench.py --workload hmc_mlp128 --no-cpu-baseline > gpurun_out/r02a_hmc.json 2> gpurun_out/r02a_hmc.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:langevin_elem -s 3 -c 1 -o gpurun_out/r02a_c2_elem -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:langevin_mlp_tc -s 3 -c 1 -o gpurun_out/r02a_mlp128_tc -f python bench.py --workload mlp128 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/r02a_tests.txt
cat gpurun_out/r02a_c2.json gpurun_out/r02a_mlp128.json gpurun_out/r02a_c3.json gpurun_out/r02a_hmc.json | cut -c1-400
