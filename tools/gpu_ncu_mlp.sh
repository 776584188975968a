# full ncu captures of the tensor-core MLP kernels that changed (wide: C3 shape through bench.py; deep: mlp128x3),
# a launch list of the default bench command, then the default bench line itself (never under the profiler)
set -x
mkdir -p gpurun_out
T=${1:-ncum}
timeout 400 ncu --set full --import-source on --clock-control none -k regex:langevin_mlp_wide_kernel -s 2 -c 1 -o gpurun_out/${T}_c3 -f python bench.py --workload c3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_c3.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:langevin_mlp_deep_kernel -s 2 -c 1 -o gpurun_out/${T}_x3 -f python bench.py --workload mlp128x3 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_x3.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:hmc_mlp_tc_kernel -s 2 -c 1 -o gpurun_out/${T}_hmc -f python bench.py --workload hmc_mlp128 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_hmc.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_c3.log gpurun_out/${T}_x3.log gpurun_out/${T}_hmc.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -n 3 gpurun_out/${T}_bench.err; cut -c1-300 gpurun_out/${T}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-300 gpurun_out/${T}_bench_ref.json
