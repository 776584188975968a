set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r02c_tests.txt
EBM_B200_STANDALONE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02c_tests_standalone.txt
python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err
cat gpurun_out/r02c_tests.txt gpurun_out/r02c_tests_standalone.txt
tail -5 gpurun_out/r02c_bench.err
python - <<'PY'
import json
l=json.load(open('gpurun_out/r02c_bench.json'))
print({k:(v if not isinstance(v,dict) else '...') for k,v in l.items()})
for k in ('e2e','native_rng','cpu_baseline','torch_cuda_baseline','triton_poc','roofline'):
    print(k, l.get(k))
for w,r in l.get('secondary',{}).items(): print(w, r)
PY
cat gpurun_out/r02c_bench_ref.json | cut -c1-300
