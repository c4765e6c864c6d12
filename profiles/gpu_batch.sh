mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_ops.py -x -q -m gpu 2>&1 | tail -1
