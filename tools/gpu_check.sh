# one-GPU round check: GPU tests, smoke(), default bench line, reference arm
set -x
mkdir -p gpurun_out
T=${1:-chk}
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/${T}_tests.txt
echo "tests_s=$SECONDS" >> gpurun_out/${T}_tests.txt
cat gpurun_out/${T}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${T}_smoke.txt
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench_done_s=$SECONDS"
tail -n 3 gpurun_out/${T}_bench.err; cut -c1-400 gpurun_out/${T}_bench.json
