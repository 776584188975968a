set -x
mkdir -p gpurun_out
T=${1:-r02i}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:langevin_mlp_wide -c 1 -o gpurun_out/${T}_c3_wide python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${T}_ncu_c3.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmc_mlp_tc -c 1 -o gpurun_out/${T}_hmc_tc python bench.py --workload hmc_mlp128 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/${T}_ncu_hmc.err
tail -2 gpurun_out/${T}_ncu_c3.err gpurun_out/${T}_ncu_hmc.err
