cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
echo "old  gap c2-shape : $(SCIPNP_LIB=$E/libscipnp_old.so timeout 200 python profiles/prof_driver.py 40 256 256 8 28 gap 2>&1 | tail -1)"
echo "new  gap c2-shape : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 gap 2>&1 | tail -1)"
echo "new nopack        : $(SCIPNP_WS_NO_PACK=1 timeout 200 python profiles/prof_driver.py 40 256 256 8 28 gap 2>&1 | tail -1)"
echo "new admm c2       : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 admm 2>&1 | tail -1)"
echo "old c1: $(SCIPNP_LIB=$E/libscipnp_old.so timeout 200 python profiles/prof_driver.py 40 256 256 8 1 gap 2>&1 | tail -1)"
echo "new c1: $(timeout 200 python profiles/prof_driver.py 40 256 256 8 1 gap 2>&1 | tail -1)"
echo "new 1024x1024x8: $(timeout 200 python profiles/prof_driver.py 40 1024 1024 8 1 gap 2>&1 | tail -1)"
echo "old 1024x1024x8: $(SCIPNP_LIB=$E/libscipnp_old.so timeout 200 python profiles/prof_driver.py 40 1024 1024 8 1 gap 2>&1 | tail -1)"
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 256 256 8 28 gap 2>&1 | grep "ws prof\|consumer  [04]\|producer  *[0-9]*:" | head -8
