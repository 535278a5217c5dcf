cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err; tail -2 gpurun_out/r24_bench.err; cut -c1-700 gpurun_out/r24_bench.json
