mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fast.py 2>&1 | tail -5 > gpurun_out/r2_pytest_exact.log
timeout 400 python -m pytest tests/test_gpu_fast.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2_pytest_fast.log
timeout 200 python tools/ab_modes.py 1000000 01 > gpurun_out/r2_ab_base.log 2>&1
for w in 1 2 8; do RTS_WALK_CTAS_PER_SM=$w timeout 200 python tools/ab_modes.py 1000000 1 >> gpurun_out/r2_ab_base.log 2>&1; done
