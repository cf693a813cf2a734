mkdir -p gpurun_out
RTS_PARITY_FRAMES=96 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 --repeats 1 --no-config4 --agents 250000 > gpurun_out/r2_par1.json 2> gpurun_out/r2_par1.err
RTS_BENCH_RANK_TRACE=1 RTS_PARITY_FRAMES=48 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 5 --repeats 5 --no-config4 > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err
