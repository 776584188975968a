set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02b_tests.txt
EBM_B200_STANDALONE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02b_tests_standalone.txt
python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r02b_c2.json 2> gpurun_out/r02b_c2.err
python bench.py --workload mlp128 --no-cpu-baseline > gpurun_out/r02b_mlp128.json 2> gpurun_out/r02b_mlp128.err
EBM_B200_LIB=$PWD/torchebm_b200/lib/libebm_b200_ld16.so python bench.py --workload mlp128 --no-cpu-baseline > gpurun_out/r02b_mlp128_ld16.json 2> gpurun_out/r02b_mlp128_ld16.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:langevin_elem -s 3 -c 1 -o gpurun_out/r02b_c2_elem -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:langevin_mlp_tc -s 3 -c 1 -o gpurun_out/r02b_mlp128_tc -f python bench.py --workload mlp128 --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
cat gpurun_out/r02b_tests.txt gpurun_out/r02b_tests_standalone.txt
cat gpurun_out/r02b_c2.json gpurun_out/r02b_mlp128.json gpurun_out/r02b_mlp128_ld16.json | cut -c1-200
