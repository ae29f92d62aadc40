#!/bin/bash
# round-2 GPU session E (N GPUs): peer-memory exchange across processes (IPC), NCCL and p2p selftests, bench at N
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2e_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k sharded_matches > $O/t_multi.log 2>&1; echo "multigpu tests rc=$?" >> $O/summary.txt
for tr in p2p nccl; do
  AMRB_TRANSPORT=$tr timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3_$tr.log 2> $O/bench_c3_$tr.err; echo "c3 $tr rc=$?" >> $O/summary.txt
  AMRB_TRANSPORT=$tr timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2_$tr.log 2> $O/bench_c2_$tr.err; echo "c2 $tr rc=$?" >> $O/summary.txt
done
tail -n 3 $O/t_multi.log; cat $O/summary.txt; grep -h -v "OMP_NUM\|\*\*\*\*\|^$" $O/*.err | tail -n 20
