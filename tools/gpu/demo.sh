mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_adapter.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2_demo.log
