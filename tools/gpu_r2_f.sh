#!/bin/bash
# round-2 GPU session F (N GPUs): the driver's scaling bench lines at N (C3 strong, p2p), C2 weak
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
O=gpurun_out/r2f_n$N; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3_p2p.log 2> $O/bench_c3_p2p.err; echo "c3 p2p rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2_p2p.log 2> $O/bench_c2_p2p.err; echo "c2 p2p rc=$?" >> $O/summary.txt
AMRB_TRANSPORT=nccl timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3_nccl.log 2> $O/bench_c3_nccl.err; echo "c3 nccl rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep -h -v "OMP_NUM\|\*\*\*\*\|^$" $O/*.err | tail -n 20
