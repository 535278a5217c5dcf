cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for e in 0 1 2; do echo "edge $e"; SCIPNP_WS_EDGE=$e timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1; done
for e in 0 1; do echo "edge $e 278"; SCIPNP_WS_EDGE=$e timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1; done
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "ws prof\] consumer-0" | head -1
timeout 900 python -m pytest tests/test_gpu_ws.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
