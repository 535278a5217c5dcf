cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for m in gap admm; do
echo "$m c2-shape pack  : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 $m 2>&1 | tail -1)"
echo "$m c2-shape nopack: $(SCIPNP_WS_NO_PACK=1 timeout 200 python profiles/prof_driver.py 40 256 256 8 28 $m 2>&1 | tail -1)"
done
echo "c3 shape: $(timeout 200 python profiles/prof_driver.py 40 256 256 24 4 gap 2>&1 | tail -1)"
echo "gap UHD: $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ws.py tests/test_gpu_configs.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
