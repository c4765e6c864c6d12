mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu -s ) > gpurun_out/r2_gpu_tests_final.log 2>&1; grep -v "^$" gpurun_out/r2_gpu_tests_final.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_final.log 2> gpurun_out/r2_bench_final.err; tail -c 1500 gpurun_out/r2_bench_final.log; tail -6 gpurun_out/r2_bench_final.err | cut -c1-600
