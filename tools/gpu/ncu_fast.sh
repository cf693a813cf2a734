mkdir -p gpurun_out
timeout 120 python tools/ncu_target.py 1 40 > gpurun_out/r2_ncu_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step --launch-skip 60 --launch-count 2 -f -o gpurun_out/r2_v1_fast python tools/ncu_target.py 1 40 > gpurun_out/r2_ncu_full.log 2>&1
