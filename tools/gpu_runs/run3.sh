cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
for e in 1 2 3 4; do
  echo "exp $e"; SCIPNP_LIB=$PWD/sci-algorithms_b200/build/exp/libscipnp_e$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gap_tv_ws -s 2 -c 1 -o gpurun_out/ws_v1 python profiles/prof_driver.py 2 > gpurun_out/ncu_ws_v1.log 2>&1
tail -3 gpurun_out/ncu_ws_v1.log
