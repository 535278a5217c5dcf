cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
E=$PWD/sci-algorithms_b200/build/exp
for e in x1 x2 x4 x8 x5 x15; do
  echo "exp $e"; SCIPNP_LIB=$E/libscipnp_$e.so timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_ws.py tests/test_dropin.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -15
for shp in "256 256 8 1 gap" "256 256 8 28 admm" "256 256 24 4 gap" "256 310 28 1 gap" "2160 3840 24 1 admm"; do timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
for g in 8 34 68 136 272; do echo "grid $g"; SCIPNP_WS_GRID=$g timeout 120 python profiles/prof_driver.py 40 256 256 8 2>&1 | tail -1; done
