cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ws.py -x -q 2>&1 | tail -30 > gpurun_out/ws_tests.log
cat gpurun_out/ws_tests.log
for v in 1 0; do
  SCIPNP_FUSED_VARIANT=$v timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
done
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | tail -60
