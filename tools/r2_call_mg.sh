set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dg_slab_parity.py 24 3 3 > gpurun_out/r2_mg3_dg24.log 2>&1; grep "rank" gpurun_out/r2_mg3_dg24.log | cut -c1-330
