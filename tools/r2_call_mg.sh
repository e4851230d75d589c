set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 50 --warmup 10 --no-dg > gpurun_out/r2_p2p_bench2.log 2>&1; tail -c 200 gpurun_out/r2_p2p_bench2.log
WB_FV2D_P2P=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 50 --warmup 10 --no-dg --no-parity > gpurun_out/r2_nccl_bench2.log 2>&1; tail -c 200 gpurun_out/r2_nccl_bench2.log
