mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-strong > gpurun_out/bench_c.log 2> gpurun_out/bench_c.err; tail -c 3000 gpurun_out/bench_c.log; tail -3 gpurun_out/bench_c.err | cut -c1-300
timeout 600 python bench.py --steps 6 --warmup 3 --no-strong --no-gpu-reference --no-cpu-baseline --clips-in-flight 2 > gpurun_out/bench_d.log 2> gpurun_out/bench_d.err; tail -c 1500 gpurun_out/bench_d.log; tail -3 gpurun_out/bench_d.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -c 1200 gpurun_out/bench_ref.log
