mkdir -p gpurun_out
RTS_BENCH_RANK_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_b8.json 2> gpurun_out/r2_b8.err
echo "rc=$?" >> gpurun_out/r2_b8.err
