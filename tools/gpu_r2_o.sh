#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2o_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3.log 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2.log 2> $O/bench_c2.err; echo "c2 rc=$?" >> $O/summary.txt
cat $O/summary.txt; grep -h -v "OMP_NUM\|\*\*\*\*\|^$\|unbatched" $O/*.err | tail -n 10
