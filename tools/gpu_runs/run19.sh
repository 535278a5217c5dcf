cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash tools/sanitize.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gap_tv_ws -s 2 -c 1 -o gpurun_out/ws_r2b python profiles/prof_driver.py 2 > gpurun_out/ncu_ws_r2b.log 2>&1
tail -2 gpurun_out/ncu_ws_r2b.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r19_bench.json 2> gpurun_out/r19_bench.err; cut -c1-600 gpurun_out/r19_bench.json
