set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_dg2d_gpu.py tests/test_fv2d_gpu.py -x -q -k "output_file or convergence" 2>&1 | tail -12 > gpurun_out/r2_c20_tests.log
cat gpurun_out/r2_c20_tests.log
