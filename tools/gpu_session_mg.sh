#!/bin/bash
# multi-GPU session (run with gpurun --gpus N): sharded-vs-single parity + the bench line at N ranks
mkdir -p gpurun_out
N=${1:-2}
( timeout 240 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q ) > gpurun_out/mg_pytest.log 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/mg_bench_n${N}.json 2> gpurun_out/mg_bench_n${N}.err
echo done
