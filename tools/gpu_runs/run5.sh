cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -5
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
SCIPNP_WS_OWN=52 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
for e in 1 2 3 4; do
  echo "exp $e"; TV_EPS=0 SCIPNP_LIB=$PWD/sci-algorithms_b200/build/exp/libscipnp_e$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "prof\] consumer-0"
timeout 300 python profiles/prof_driver.py 20 286 3840 24 2>&1 | tail -1
