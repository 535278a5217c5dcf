cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
echo "main (f_old from smem)"; timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
echo "tmaw (dedicated TMA warp)"; SCIPNP_LIB=$E/libscipnp_tmaw.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -5
SCIPNP_LIB=$E/libscipnp_tmaw.so timeout 900 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -5
SCIPNP_LIB=$E/libscipnp_tmaw.so SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "ws prof\|consumer\|producer" | head -18
for shp in "256 256 8 1 gap" "256 256 8 28 admm" "256 256 24 4 gap" "256 310 28 1 gap" "286 3840 24 1 gap"; do SCIPNP_LIB=$E/libscipnp_tmaw.so timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
