set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -E "own-norm|passed|failed|Error|error" | tail -40 > gpurun_out/r2_c13_tests.log
cat gpurun_out/r2_c13_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 900 python bench.py ) > gpurun_out/r2_c13_bench.json 2> gpurun_out/r2_c13_bench.err; tail -c 6000 gpurun_out/r2_c13_bench.json; tail -5 gpurun_out/r2_c13_bench.err
