# wide-kernel iteration: parity tests that reach langevin_mlp_wide_kernel, timeline of CTA 0, C3 timing
set -x
mkdir -p gpurun_out
T=${1:-wd}
timeout 600 python -m pytest tests/test_gpu_langevin.py tests/test_gpu_pcd.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${T}_tests.txt
[ -f torchebm_b200/lib/libebm_b200_trace.so ] && EBM_B200_LIB=torchebm_b200/lib/libebm_b200_trace.so timeout 200 python tools/wd_trace.py 784 20 512 > gpurun_out/${T}_trace.txt 2> gpurun_out/${T}_trace.err
timeout 200 python tools/mlp_probe.py 784 20 148,512 2>&1 | tee gpurun_out/${T}_probe.txt
timeout 300 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/${T}_c3.json 2> gpurun_out/${T}_c3.err; tail -2 gpurun_out/${T}_c3.err; cut -c1-200 gpurun_out/${T}_c3.json
