set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dg_slab_parity.py 64 3 2 > gpurun_out/r2_dgp2p_a.log 2>&1; grep -E "rank 0|kinds|Error" gpurun_out/r2_dgp2p_a.log | cut -c1-260 | head -20
