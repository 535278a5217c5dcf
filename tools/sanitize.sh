#!/bin/bash
# compute-sanitizer over the fused kernels (run on the GPU box): logs -> gpurun_out/sanitizer_*.log
# memcheck / synccheck / racecheck on the production library; racecheck a second time on the WS_SANITIZE=1 build
# (tools/build_exp.sh sanitize "-DSCIPNP_FUSED_FAST_BUILD -DWS_SANITIZE=1"): there every lane arrives on the
# mbarriers itself, so the tool can follow the producer -> consumer hand-over thread by thread (with one elected
# lane behind a __syncwarp it reports the other 31 lanes' accesses as unordered).
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_driver.py all > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max\||mean shift|launches" gpurun_out/sanitizer_$tool.log | tail -8
done
L=$PWD/sci-algorithms_b200/build/exp/libscipnp_sanitize.so
if [ -f $L ]; then
  SCIPNP_LIB=$L timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_driver.py all > gpurun_out/sanitizer_racecheck_all_lanes_arrive.log 2>&1
  echo "== racecheck (all lanes arrive): rc=$?"; grep -E "RACECHECK SUMMARY|max\||mean shift|launches" gpurun_out/sanitizer_racecheck_all_lanes_arrive.log | tail -8
fi
