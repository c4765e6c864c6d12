mkdir -p gpurun_out
PROPAINTER_B200_LIB=$PWD/propainter_b200/libpropainter_b200_prof.so timeout 200 python profiles/conv_prof.py > gpurun_out/conv_prof5.log 2>&1
grep -A4 "3x3 gen 128->128 bn64\|3x3 rfc 128->128 auto" gpurun_out/conv_prof5.log
timeout 300 python profiles/conv_check.py step > gpurun_out/conv_step5.log 2>&1; cat gpurun_out/conv_step5.log
timeout 300 python profiles/conv_check.py basic > gpurun_out/conv_basic5.log 2>&1; cat gpurun_out/conv_basic5.log
timeout 200 python profiles/bisect_c2.py default calls=4 > gpurun_out/bisect2.log 2>&1
timeout 200 python profiles/bisect_c2.py cudnn umma=0 calls=4 >> gpurun_out/bisect2.log 2>&1
grep call gpurun_out/bisect2.log
