mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 600 python bench.py --no-cpu-baseline ) > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; tail -c 2500 gpurun_out/bench_1gpu.log; tail -4 gpurun_out/bench_1gpu.err | cut -c1-300
timeout 600 $TR profiles/dist_check.py 33 10 > gpurun_out/dist_check_2gpu.log 2>&1; grep -v "^W\|warn" gpurun_out/dist_check_2gpu.log | tail -6
( time timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; tail -c 3500 gpurun_out/bench_2gpu.log; tail -4 gpurun_out/bench_2gpu.err | cut -c1-300
