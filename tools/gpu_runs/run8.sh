cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
TV_EPS=0 SCIPNP_LIB=$PWD/sci-algorithms_b200/build/exp/libscipnp_x4.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
timeout 300 python profiles/prof_driver.py 20 286 3840 24 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
