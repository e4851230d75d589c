set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c70_tests.log 2>&1; tail -3 gpurun_out/r2_c70_tests.log
timeout 500 python bench.py > gpurun_out/r2_c70_bench.json 2> gpurun_out/r2_c70_bench.err; tail -c 200 gpurun_out/r2_c70_bench.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_c70_smoke.log 2>&1; tail -2 gpurun_out/r2_c70_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-big > gpurun_out/r2_final_launch_bench.log 2>&1; tail -c 200 gpurun_out/r2_final_launch_bench.log
