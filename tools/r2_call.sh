set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dg2d_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/r2_c5_tests.log
cat gpurun_out/r2_c5_tests.log
for n in 4096 8192; do timeout 300 python tools/dg2d_rate.py $n 3 4; done 2>&1 | grep "^DG" | tee gpurun_out/r2_c5_rates.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dg_stage_split -s 10 -c 5 -o gpurun_out/r2_split_d python tools/dg2d_rate.py 4096 3 2 > gpurun_out/r2_c5_ncu.log 2>&1
tail -3 gpurun_out/r2_c5_ncu.log
