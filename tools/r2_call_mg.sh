set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/slab_parity.py 96 70 6 > gpurun_out/r2_p2p_a.log 2>&1; grep -E "rank 0|kinds|Error|error" gpurun_out/r2_p2p_a.log | cut -c1-250 | head -20
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/slab_parity.py 130 257 4 > gpurun_out/r2_p2p_b.log 2>&1; grep -E "rank 0|kinds|Error|error" gpurun_out/r2_p2p_b.log | cut -c1-250 | head -20
