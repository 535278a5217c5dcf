cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
for lib in main p3off; do
  if [ $lib = main ]; then unset SCIPNP_LIB; else export SCIPNP_LIB=$E/libscipnp_$lib.so; fi
  for shp in "2160 3840 24" "278 3840 24"; do echo "$lib $shp"; timeout 200 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
done
unset SCIPNP_LIB
for c in 4 8 12; do echo "segcost $c"; SCIPNP_WS_SEGCOST=$c timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1; done
for shp in "256 256 8 1 gap" "256 256 24 4 gap"; do timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_gpu_ws.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -5
