cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TV_EPS=0 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
for e in 1 2 3 4; do
  echo "exp $e"; TV_EPS=0 SCIPNP_LIB=$PWD/sci-algorithms_b200/build/exp/libscipnp_e$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
