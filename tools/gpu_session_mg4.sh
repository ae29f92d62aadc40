#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/mg_bench_n${N}.json 2> gpurun_out/mg_bench_n${N}.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29732 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/mg_ref_n${N}.json 2> gpurun_out/mg_ref_n${N}.err
echo done
