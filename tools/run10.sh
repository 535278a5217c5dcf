cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "ws prof\|consumer\|producer" | head -60
for e in x1 x2 x4 x8 x15; do
  echo "exp $e"; SCIPNP_LIB=$E/libscipnp_$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
