# one full ncu capture (with source) of langevin_mlp_wide_kernel at the C3 shape; the report is read here afterwards
set -x
mkdir -p gpurun_out
T=${1:-ncuw}
timeout 500 ncu --set full --import-source on --clock-control none -k regex:langevin_mlp_wide_kernel -s 3 -c 1 -o gpurun_out/${T} -f python tools/mlp_probe.py 784 20 512 > gpurun_out/${T}.log 2>&1
tail -3 gpurun_out/${T}.log
ls -la gpurun_out/${T}.ncu-rep
