cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for e in 0 4 7 10; do echo "edge $e"; SCIPNP_WS_EDGE=$e timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1; done
for e in 0 7; do echo "edge $e 278"; SCIPNP_WS_EDGE=$e timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1; done
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "ws prof\] consumer-0" | head -2
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 256 256 8 2>&1 | grep "ws prof\|consumer  0\|producer 1[35]" | head -12
for g in 33 66 132; do echo "c1 grid $g"; SCIPNP_WS_GRID=$g timeout 120 python profiles/prof_driver.py 40 256 256 8 2>&1 | tail -1; done
bash tools/sanitize.sh 2>&1 | grep -A6 "all lanes"
timeout 900 python -m pytest tests/test_gpu_ws.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
