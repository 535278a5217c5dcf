cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "admm c2 ws    : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
echo "admm c2 stream: $(SCIPNP_FUSED_VARIANT=1 timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
echo "admm 512x512x16 ws    : $(timeout 200 python profiles/prof_driver.py 40 512 512 16 4 admm 2>&1 | tail -1)"
echo "admm 512x512x16 stream: $(SCIPNP_FUSED_VARIANT=1 timeout 200 python profiles/prof_driver.py 40 512 512 16 4 admm 2>&1 | tail -1)"
echo "admm UHD ws: $(timeout 200 python profiles/prof_driver.py 40 2160 3840 24 1 admm 2>&1 | tail -1)"
echo "gap UHD: $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ws.py tests/test_gpu_configs.py -x -q 2>&1 | tail -3
