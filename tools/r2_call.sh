set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dg2d_gpu.py tests/test_reference_pins_gpu.py -q -s 2>&1 | grep -E "^FAILED|passed|failed|HIO fused" | tail -30 > gpurun_out/r2_c21_tests.log
cat gpurun_out/r2_c21_tests.log
