cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python tools/time_load.py 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "one_pass_init or golden or operators" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu > gpurun_out/r32_bench.json 2> gpurun_out/r32_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r32_bench.json'))
print(d['value'], d['ms_per_step'], d['ms_per_iteration'], d['roofline']['traffic'], d['roofline']['traffic_source'])"
