set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c50_tests.log 2>&1; tail -3 gpurun_out/r2_c50_tests.log
timeout 500 python bench.py > gpurun_out/r2_c50_bench.json 2> gpurun_out/r2_c50_bench.err; tail -c 300 gpurun_out/r2_c50_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c50_ref.json 2> gpurun_out/r2_c50_ref.err; tail -c 600 gpurun_out/r2_c50_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_c50_smoke.log 2>&1; tail -2 gpurun_out/r2_c50_smoke.log
