set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_stage_tma -s 8 -c 2 -o gpurun_out/r2_fv_a python bench.py --steps 3 --warmup 3 --no-dg --no-cpu --no-e2e > gpurun_out/r2_c12_ncu.log 2>&1
tail -2 gpurun_out/r2_c12_ncu.log | cut -c1-300
