cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo main; timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
for e in base rcp1 rcp2 regs regs136 rcp1regs rcp1regs136 rcp1x4; do
  echo "exp $e"; SCIPNP_LIB=$E/libscipnp_$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
echo "rcp1 tests"; SCIPNP_LIB=$E/libscipnp_rcp1.so timeout 600 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -3
echo "rcp1regs136 tests"; SCIPNP_LIB=$E/libscipnp_rcp1regs136.so timeout 600 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gap_tv_ws -s 2 -c 1 -o gpurun_out/ws_r2a python profiles/prof_driver.py 2 > gpurun_out/ncu_ws_r2a.log 2>&1
tail -3 gpurun_out/ncu_ws_r2a.log
