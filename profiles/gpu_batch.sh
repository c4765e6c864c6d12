mkdir -p gpurun_out
timeout 120 profiles/probes/umma_rate_probe > gpurun_out/umma_rate.log 2>&1
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 300 python profiles/conv_prof.py > gpurun_out/conv_prof3.log 2>&1
timeout 900 python profiles/conv_check.py > gpurun_out/conv_check3.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -x -q -m gpu -s > gpurun_out/pytest_ops_modules.log 2>&1
tail -5 gpurun_out/pytest_ops_modules.log
