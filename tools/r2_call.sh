set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 $NCU -k regex:k_stage_tma -s 12 -c 2 -o gpurun_out/r2_final_fv python tools/fv2d_perf.py 4096 4 > gpurun_out/r2_final_fv.log 2>&1; tail -2 gpurun_out/r2_final_fv.log
python profiles/ncu_summary.py gpurun_out/r2_final_fv.ncu-rep 16777216 > gpurun_out/r2_final_fv_summary.txt 2>&1
timeout 500 $NCU -k regex:k_dg_stage_split -s 10 -c 5 -o gpurun_out/r2_final_dg python tools/dg2d_rate.py 4096 3 2 > gpurun_out/r2_final_dg.log 2>&1; tail -2 gpurun_out/r2_final_dg.log
python profiles/ncu_summary.py gpurun_out/r2_final_dg.ncu-rep 16777216 > gpurun_out/r2_final_dg_summary.txt 2>&1
timeout 500 $NCU -k "regex:k_dg_stage_split|k_limiter_hio_onp" -s 20 -c 4 -o gpurun_out/r2_final_dg_hio python tools/dg2d_rate.py 2048 3 2 HIO > gpurun_out/r2_final_dg_hio.log 2>&1; tail -2 gpurun_out/r2_final_dg_hio.log
python profiles/ncu_summary.py gpurun_out/r2_final_dg_hio.ncu-rep 4194304 > gpurun_out/r2_final_dg_hio_summary.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launch_list.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-big > gpurun_out/r2_final_launch_bench.log 2>&1; tail -c 300 gpurun_out/r2_final_launch_bench.log
du -sm gpurun_out; ls -la gpurun_out/
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/r2_final_dg_hio.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 60 ]; then rm -f gpurun_out/r2_final_fv.ncu-rep; fi
