mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_all.log
timeout 500 python tools/ab_config4.py 10000000 6 0 > gpurun_out/r2_c4_cs.log 2>&1
timeout 200 python tools/ab_modes.py 1000000 1 >> gpurun_out/r2_c4_cs.log 2>&1
