cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
for lib in main genu; do
  if [ $lib = main ]; then unset SCIPNP_LIB; else export SCIPNP_LIB=$E/libscipnp_$lib.so; fi
  for shp in "2160 3840 24" "278 3840 24"; do echo "$lib $shp"; timeout 200 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
done
unset SCIPNP_LIB
for c in 0 8 24 32 48; do echo "segcost $c"; SCIPNP_WS_SEGCOST=$c timeout 200 python profiles/prof_driver.py 40 278 3840 24 2>&1 | tail -1; done
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 278 3840 24 2>&1 | grep "ws prof\|consumer  0\|producer 13" | head -12
