cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
for e in x1 x2 x4 x8 x15 rcp1 rcp2 cr152; do
  echo "exp $e"; SCIPNP_LIB=$E/libscipnp_$e.so TV_EPS=0 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
