cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${NGPU:-2}
KS=1,2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/tiled_timing.py 2>&1 | grep "N=" 
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r16_bench_n$N.json 2> gpurun_out/r16_bench_n$N.err; tail -3 gpurun_out/r16_bench_n$N.err; cat gpurun_out/r16_bench_n$N.json
