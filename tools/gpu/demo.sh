mkdir -p gpurun_out
timeout 300 oracle/_ref/adapter_demo 2000 40 80 7 > gpurun_out/r2_demo.log 2>&1
timeout 400 python -m pytest tests/test_gpu_adapter.py -m gpu -q 2>&1 | tail -5 >> gpurun_out/r2_demo.log
