cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
echo base; TV_EPS=0 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
echo swz; SCIPNP_WS_SWZ=1 timeout 600 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -3
SCIPNP_WS_SWZ=1 TV_EPS=0 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
for e in rcp2 rcp2x1 x8 rcp2x8 rcp2x15; do
  echo "exp $e"; TV_EPS=0 SCIPNP_LIB=$E/libscipnp_$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
echo "rcp2 tests"; SCIPNP_LIB=$E/libscipnp_rcp2.so timeout 600 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -3
echo "rcp2+swz"; SCIPNP_WS_SWZ=1 TV_EPS=0 SCIPNP_LIB=$E/libscipnp_rcp2.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
