set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fv2d_multigpu.py -x -q 2>&1 | tail -5 > gpurun_out/r2_mg_tests.log; cat gpurun_out/r2_mg_tests.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-dg --no-cpu --no-e2e > gpurun_out/r2_mg_bench_n1.json 2>gpurun_out/r2_mg_bench_n1.err; tail -c 700 gpurun_out/r2_mg_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_mg_bench_n2.json 2> gpurun_out/r2_mg_bench_n2.err; tail -c 3000 gpurun_out/r2_mg_bench_n2.json; tail -5 gpurun_out/r2_mg_bench_n2.err
