cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for shp in "256 256 8 1 gap" "256 256 8 28 gap" "512 512 8 4 gap" "1024 1024 8 1 gap" "256 256 4 8 gap"; do echo "$shp: $(timeout 200 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1)"; done
echo "admm c2 (stream): $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 256 256 8 28 gap 2>&1 | grep "ws prof\|consumer  [04]\|producer  *[0-9]*:" | head -8
timeout 1500 python -m pytest tests/test_gpu_ws.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
