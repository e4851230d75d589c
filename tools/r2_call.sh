set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_dg2d_gpu.py tests/test_reference_pins_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2_c15_tests.log
cat gpurun_out/r2_c15_tests.log
( for n in 4096 8192; do timeout 200 python tools/dg2d_rate.py $n 3 4; done; timeout 200 python tools/dg2d_rate.py 8192 3 10 ) 2>&1 | grep "^DG" | tee gpurun_out/r2_c15_rates.log
