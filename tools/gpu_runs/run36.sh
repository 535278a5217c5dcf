cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for m in gap admm; do
echo "$m c2-shape pack  : $(timeout 200 python profiles/prof_driver.py 40 256 256 8 28 $m 2>&1 | tail -1)"
echo "$m c2-shape nopack: $(SCIPNP_WS_NO_PACK=1 timeout 200 python profiles/prof_driver.py 40 256 256 8 28 $m 2>&1 | tail -1)"
done
echo "c1: $(timeout 200 python profiles/prof_driver.py 40 256 256 8 1 gap 2>&1 | tail -1)"
echo "gap UHD: $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 256 256 8 28 gap 2>&1 | grep "ws prof\|consumer  [04]\|producer  *[0-9]*:" | head -8
