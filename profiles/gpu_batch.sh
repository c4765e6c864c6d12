mkdir -p gpurun_out
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 200 python profiles/conv_prof.py > gpurun_out/conv_prof7.log 2>&1
grep -A4 "3x3 gen 128->128 bn64\|3x3 rfc 128->128 auto\|1x1 gen K=1152" gpurun_out/conv_prof7.log
timeout 300 python profiles/conv_check.py step > gpurun_out/conv_step7.log 2>&1; cat gpurun_out/conv_step7.log
timeout 300 python profiles/conv_check.py basic > gpurun_out/conv_basic7.log 2>&1; grep "own" gpurun_out/conv_basic7.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu > gpurun_out/pytest_ops7.log 2>&1; tail -3 gpurun_out/pytest_ops7.log
PP_UMMA_CONV=1 timeout 200 python profiles/bisect_c2.py umma calls=4 > gpurun_out/bisect3.log 2>&1
PP_UMMA_CONV=0 timeout 200 python profiles/bisect_c2.py cudnn calls=4 >> gpurun_out/bisect3.log 2>&1
grep call gpurun_out/bisect3.log
