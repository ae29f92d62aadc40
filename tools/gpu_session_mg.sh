#!/bin/bash
# multi-GPU session (run with gpurun --gpus N): sharded-vs-single parity + the bench line at N ranks
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/mg_smi.txt 2>&1
( timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q ) > gpurun_out/mg_pytest.log 2>&1
for g in 1 0; do
AMRB_GRAPH=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$g bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/mg_bench_n${N}_graph$g.json 2> gpurun_out/mg_bench_n${N}_graph$g.err
done
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/mg_bench_n1.json 2> gpurun_out/mg_bench_n1.err
echo done
