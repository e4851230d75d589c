set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fv2d_multigpu.py -x -q -m gpu > gpurun_out/r2_mg6_tests.log 2>&1; tail -3 gpurun_out/r2_mg6_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_mg6_bench2.log 2>&1; tail -c 200 gpurun_out/r2_mg6_bench2.log
