#!/bin/bash
# round-2 GPU session B (N GPUs): NCCL parity selftests, strong-scaling C3 bench at N, weak C2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2b_n$N; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x > $O/t_multi.log 2>&1; echo "multigpu tests rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c3L4 --steps 10 --warmup 3 > $O/bench_c3L4.log 2> $O/bench_c3L4.err; echo "c3L4 rc=$?" >> $O/summary.txt
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_c3.log 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2.log 2> $O/bench_c2.err; echo "c2 rc=$?" >> $O/summary.txt
tail -n 3 $O/t_multi.log; cat $O/summary.txt; tail -n 5 $O/*.err
