cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r12_tests.log
cat gpurun_out/r12_tests.log
timeout 600 python bench.py > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err; tail -3 gpurun_out/r12_bench.err; cat gpurun_out/r12_bench.json
timeout 300 python profiles/prof_driver.py 20 2>&1 | tail -1
SCIPNP_WS_PROF=1 timeout 300 python profiles/prof_driver.py 6 2>&1 | grep "ws prof\|consumer\|producer" | head -18
for shp in "256 256 8 1 gap" "256 256 8 28 admm" "256 256 24 4 gap" "256 310 28 1 gap" "286 3840 24 1 gap" "2160 3840 24 1 admm"; do timeout 120 python profiles/prof_driver.py 40 $shp 2>&1 | tail -1; done
