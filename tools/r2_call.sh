set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c60_tests.log 2>&1; tail -3 gpurun_out/r2_c60_tests.log
timeout 500 python bench.py > gpurun_out/r2_c60_bench.json 2> gpurun_out/r2_c60_bench.err; tail -c 200 gpurun_out/r2_c60_bench.json
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 500 $NCU -k regex:k_dg_stage_split -s 10 -c 5 -o gpurun_out/r2_final_dg_src python tools/dg2d_rate.py 4096 3 2 ONP 2 2 1 1 > gpurun_out/r2_final_dg_src.log 2>&1; tail -1 gpurun_out/r2_final_dg_src.log
python profiles/ncu_summary.py gpurun_out/r2_final_dg_src.ncu-rep 16777216 > gpurun_out/r2_final_dg_src_summary.txt 2>&1
python profiles/ncu_hot.py gpurun_out/r2_final_dg_src.ncu-rep 1 128 >> gpurun_out/r2_final_dg_src_summary.txt 2>&1
rm -f gpurun_out/r2_final_dg_src.ncu-rep
