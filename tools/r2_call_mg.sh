set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --steps 50 --warmup 10 > gpurun_out/r2_bench_n8c.log 2>&1; tail -c 200 gpurun_out/r2_bench_n8c.log
