set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_fv2d_gpu.py -x -q -m gpu > gpurun_out/r2_c30_tests.log 2>&1; tail -3 gpurun_out/r2_c30_tests.log
timeout 300 python bench.py --no-dg --no-cpu --no-e2e > gpurun_out/r2_c30_bench.json 2> gpurun_out/r2_c30_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r2_c30_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['clocks'], d.get('fv2d_16384_n1'))"
