set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c31_tests.log 2>&1; tail -3 gpurun_out/r2_c31_tests.log
timeout 400 python bench.py > gpurun_out/r2_c31_bench.json 2> gpurun_out/r2_c31_bench.err; tail -c 600 gpurun_out/r2_c31_bench.json
