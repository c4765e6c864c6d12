mkdir -p gpurun_out
timeout 120 profiles/probes/umma_rate_probe > gpurun_out/umma_rate.log 2>&1
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 300 python profiles/conv_prof.py > gpurun_out/conv_prof3.log 2>&1
timeout 900 python profiles/conv_check.py > gpurun_out/conv_check3.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -q -m gpu -s > gpurun_out/pytest_ops_modules.log 2>&1
tail -5 gpurun_out/pytest_ops_modules.log
timeout 300 python profiles/check_experiments.py > gpurun_out/exp_attn_default.log 2>&1
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_uloop.so timeout 300 python profiles/check_experiments.py > gpurun_out/exp_attn_uloop.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_zz_full_size.py -q -m gpu -s > gpurun_out/pytest_full.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-strong > gpurun_out/bench_a.log 2> gpurun_out/bench_a.err
timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-strong --no-gpu-reference --clips-in-flight 2 > gpurun_out/bench_b.log 2> gpurun_out/bench_b.err
tail -3 gpurun_out/bench_a.log gpurun_out/bench_b.log
PP_PDL=0 timeout 600 python -m pytest tests/test_gpu_modules.py -q -m gpu -k "flow_completion or generator_matches" > gpurun_out/pytest_nopdl.log 2>&1
timeout 300 python profiles/ncu_targets.py --time > gpurun_out/kernel_times_r2.txt 2>&1
