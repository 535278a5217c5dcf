cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "gap UHD: $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
echo "gap 278: $(timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1)"
echo "c3 shape: $(timeout 200 python profiles/prof_driver.py 40 256 256 24 4 gap 2>&1 | tail -1)"
timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ws.py tests/test_gpu_configs.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
