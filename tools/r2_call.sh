set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dg2d_gpu.py tests/test_reference_pins_gpu.py -q 2>&1 | grep -E "^FAILED|passed|failed" | tail -30 > gpurun_out/r2_c26_tests.log
cat gpurun_out/r2_c26_tests.log
