set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -30 > gpurun_out/ws_tests.log
cat gpurun_out/ws_tests.log
for v in 1 0; do
  SCIPNP_FUSED_VARIANT=$v timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -2
done
SCIPNP_WS_OWN=52 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
SCIPNP_WS_OWN=48 SCIPNP_WS_NSEG=2 timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
timeout 300 python profiles/prof_driver.py 20 286 3840 24 2>&1 | tail -1
SCIPNP_FUSED_VARIANT=1 timeout 300 python profiles/prof_driver.py 20 286 3840 24 2>&1 | tail -1
for shp in "256 256 8" "512 512 24" "256 320 24"; do
  for v in 1 0; do SCIPNP_FUSED_VARIANT=$v timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
done
