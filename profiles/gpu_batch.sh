mkdir -p gpurun_out
PROF_QUICK=1 PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 100 python profiles/conv_prof.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_modules.py -q -m gpu -s > gpurun_out/pytest_om8.log 2>&1; grep -v "^$" gpurun_out/pytest_om8.log | tail -25
timeout 900 python -m pytest tests/test_gpu_zz_full_size.py -q -m gpu -s > gpurun_out/pytest_full8.log 2>&1; grep -v "^$" gpurun_out/pytest_full8.log | tail -12
