cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for shp in "2160 3840 24 1 gap" "278 3840 24 1 gap" "256 256 8 1 gap" "256 256 24 4 gap"; do
  echo "pdl   $shp: $(timeout 200 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1)"
  echo "nopdl $shp: $(SCIPNP_WS_NO_PDL=1 timeout 200 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1)"
done
timeout 1200 python -m pytest tests/test_gpu_ws.py tests/test_gpu_tiled.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "rec_loops or admm_denoise_bayer" 2>&1 | tail -3
