cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "admm UHD ws    : $(timeout 200 python profiles/prof_driver.py 40 2160 3840 24 1 admm 2>&1 | tail -1)"
echo "admm c2 ws    : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
echo "admm c2 stream: $(SCIPNP_FUSED_VARIANT=1 timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
echo "gap  c2-shape ws: $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 gap 2>&1 | tail -1)"
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 256 256 8 28 admm 2>&1 | grep "ws prof\|consumer  [04]\|producer 1[35]" | head -9
