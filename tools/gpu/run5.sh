mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neighbour_queries or alignment or incremental" 2>&1 | tail -15 > gpurun_out/r2_pytest_new.log
