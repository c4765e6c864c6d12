mkdir -p gpurun_out
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 200 python profiles/conv_prof.py > gpurun_out/conv_prof6.log 2>&1
grep -A4 "3x3 gen 128->128 bn64\|3x3 rfc 128->128 auto\|1x1 gen K=1152" gpurun_out/conv_prof6.log
timeout 300 python profiles/conv_check.py step > gpurun_out/conv_step6.log 2>&1; cat gpurun_out/conv_step6.log
timeout 300 python profiles/conv_check.py deform > gpurun_out/conv_deform6.log 2>&1; cat gpurun_out/conv_deform6.log
timeout 400 python profiles/conv_check.py shapes > gpurun_out/conv_shapes6.log 2>&1; grep "1x1\|Cout=432\|ragged" gpurun_out/conv_shapes6.log
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu > gpurun_out/pytest_ops6.log 2>&1; tail -3 gpurun_out/pytest_ops6.log
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_uloop.so timeout 200 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "attention" > gpurun_out/pytest_uloop.log 2>&1; tail -2 gpurun_out/pytest_uloop.log
