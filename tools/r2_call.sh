set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dg2d_gpu.py tests/test_reference_pins_gpu.py -x -q -m gpu > gpurun_out/r2_c42_tests.log 2>&1; tail -3 gpurun_out/r2_c42_tests.log
( timeout 200 python tools/dg2d_rate.py 4096 3 4 ONP 2 2 1 2 2>&1 | tail -1
WB_DG2D_GSEP=0 timeout 200 python tools/dg2d_rate.py 4096 3 4 ONP 2 2 1 2 2>&1 | tail -1
timeout 200 python tools/dg2d_rate.py 4096 3 4 ONP 2 2 2 2 2>&1 | tail -1
timeout 200 python tools/dg2d_rate.py 8192 3 4 ONP 2 2 1 1 2>&1 | tail -1 ) > gpurun_out/r2_c42_rates.log 2>&1
cat gpurun_out/r2_c42_rates.log
