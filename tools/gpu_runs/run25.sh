cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for i in 1 2; do
echo "copy   : $(timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
echo "borrow : $(BORROW=1 timeout 200 python profiles/prof_driver.py 40 2>&1 | tail -1)"
done
timeout 600 python bench.py --no-cpu > gpurun_out/r25_bench.json 2> gpurun_out/r25_bench.err; python -c "
import json
d=json.load(open('gpurun_out/r25_bench.json'))
print(d['value'], d['ms_per_step'], d['ms_per_iteration'], d['clocks'])"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
