cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -4
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
SCIPNP_WS_OWN=52 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "prof\] consumer-0"
timeout 300 python profiles/prof_driver.py 20 286 3840 24 2>&1 | tail -1
timeout 300 python profiles/prof_driver.py 20 556 3840 24 2>&1 | tail -1
for shp in "256 256 8" "512 512 24" "256 320 24"; do timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
