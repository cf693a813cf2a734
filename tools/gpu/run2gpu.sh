mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
echo "rc=$?" >> gpurun_out/r2_b2.err
