mkdir -p gpurun_out
python tools/ab_config5.py > gpurun_out/r2_c5_var.log 2>&1
python tools/ab_modes.py 1000000 1 >> gpurun_out/r2_c5_var.log 2>&1
for so in rtsengine_b200/variants/librts_*.so; do
  RTS_B200_LIB=$PWD/$so timeout 300 python tools/ab_config5.py 2>&1 | tail -1 >> gpurun_out/r2_c5_var.log
  RTS_B200_LIB=$PWD/$so timeout 300 python tools/ab_modes.py 1000000 1 2>&1 | tail -1 >> gpurun_out/r2_c5_var.log
done
