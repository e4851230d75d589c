set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fv2d_multigpu.py -x -q -m gpu > gpurun_out/r2_mg4_tests.log 2>&1; tail -5 gpurun_out/r2_mg4_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_mg4_bench2.log 2>&1; tail -2 gpurun_out/r2_mg4_bench2.log
