set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/r2_bench_n8b.log 2>&1; tail -c 300 gpurun_out/r2_bench_n8b.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 tools/slab_parity.py 130 257 4 > gpurun_out/r2_p2p_n4.log 2>&1; grep -E "rank 0|kinds" gpurun_out/r2_p2p_n4.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29550 tools/dg_slab_parity.py 24 3 3 > gpurun_out/r2_dg_n4.log 2>&1; grep -E "rank 0" gpurun_out/r2_dg_n4.log | cut -c1-200
