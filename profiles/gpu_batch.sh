mkdir -p gpurun_out
timeout 400 python profiles/host_time2.py > gpurun_out/host_time2.log 2>&1; tail -12 gpurun_out/host_time2.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_c.log 2> gpurun_out/bench_c.err; tail -c 2500 gpurun_out/bench_c.log; tail -3 gpurun_out/bench_c.err | cut -c1-300
timeout 600 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline --clips-in-flight 2 > gpurun_out/bench_d.log 2> gpurun_out/bench_d.err; tail -c 1500 gpurun_out/bench_d.log; tail -3 gpurun_out/bench_d.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.log; tail -2 gpurun_out/bench_ref.err | cut -c1-300
