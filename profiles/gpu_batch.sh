mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/pytest_ops.log 2>&1; tail -3 gpurun_out/pytest_ops.log
timeout 200 python profiles/bisect_c2.py default > gpurun_out/bisect.log 2>&1
PP_PDL=0 timeout 200 python profiles/bisect_c2.py nopdl >> gpurun_out/bisect.log 2>&1
timeout 200 python profiles/bisect_c2.py cudnn umma=0 >> gpurun_out/bisect.log 2>&1
timeout 200 python profiles/bisect_c2.py wif1 wif=1 >> gpurun_out/bisect.log 2>&1
timeout 200 python profiles/bisect_c2.py attnmma attn=mma >> gpurun_out/bisect.log 2>&1
cat gpurun_out/bisect.log | grep "call"
