cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tiled.py -x -q 2>&1 | tail -25 > gpurun_out/r15_tests.log
cat gpurun_out/r15_tests.log
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
