cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
bash tools/sanitize.sh
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r42_tests.log; cat gpurun_out/r42_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gap_tv_ws -s 2 -c 1 -o gpurun_out/ws_r2d python profiles/prof_driver.py 2 > gpurun_out/ncu_ws_r2d.log 2>&1; tail -1 gpurun_out/ncu_ws_r2d.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 python bench.py > gpurun_out/r42_bench.json 2> gpurun_out/r42_bench.err; cut -c1-200 gpurun_out/r42_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r42_ref.json 2>/dev/null; cut -c1-200 gpurun_out/r42_ref.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
