set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for r in 32 64 128 256; do WB_DG2D_ROWS=$r timeout 200 python tools/dg2d_rate.py 8192 3 4 ONP 2>&1 | tail -1; done
for v in 1 2 3 4; do WB_DG2D_SCHED=$v timeout 200 python tools/dg2d_rate.py 8192 3 4 ONP 2>&1 | tail -1; done ) > gpurun_out/r2_c61_rates.log 2>&1
cut -c1-150 gpurun_out/r2_c61_rates.log
