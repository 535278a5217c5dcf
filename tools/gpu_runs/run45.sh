cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_iteration'], d['roofline']['traffic'], d['roofline']['traffic_source'][:60])"
