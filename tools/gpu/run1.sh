mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fast.py 2>&1 | tail -15 > gpurun_out/r2_pytest_exact.log
timeout 400 python -m pytest tests/test_gpu_fast.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/r2_pytest_fast.log
timeout 200 python tools/ab_modes.py 1000000 01 > gpurun_out/r2_ab_base.log 2>&1
bash tools/run_variants.sh 1 > gpurun_out/r2_ab_variants.log 2>&1
