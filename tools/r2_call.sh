set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dg2d_gpu.py -x -q -s -k "twelve or convergence" 2>&1 | grep -E "order|passed|failed|Error|assert" | tail -20 > gpurun_out/r2_c16_tests.log
cat gpurun_out/r2_c16_tests.log
