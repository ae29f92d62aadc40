#!/bin/bash
# round-2 GPU session N (8 GPUs): final scaling lines (C3 strong: schedule probe incl. overlap; C2 weak), p2p
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out/r2n_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3.log 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/summary.txt
timeout 400 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2.log 2> $O/bench_c2.err; echo "c2 rc=$?" >> $O/summary.txt
timeout 400 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k "sharded_matches and p2p" > $O/t_multi.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 3 $O/t_multi.log; grep -h -v "OMP_NUM\|\*\*\*\*\|^$\|unbatched" $O/*.err | tail -n 10
