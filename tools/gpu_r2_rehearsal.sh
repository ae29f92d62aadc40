#!/bin/bash
# driver rehearsal at N GPUs: reference arm, our arm, exactly the driver's command lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-1}
O=gpurun_out/r2r_n$N; mkdir -p $O
if [ "$N" = "1" ]; then
  ( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/ref.log 2> $O/ref.err; echo "ref rc=$?" >> $O/summary.txt
  ( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/ours.log 2> $O/ours.err; echo "ours rc=$?" >> $O/summary.txt
  ( time python3 -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
  ( time $TR bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > $O/ref.log 2> $O/ref.err; echo "ref rc=$?" >> $O/summary.txt
  ( time $TR bench.py --gpus $N --steps 20 --warmup 5 ) > $O/ours.log 2> $O/ours.err; echo "ours rc=$?" >> $O/summary.txt
  ( time $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 ) > $O/c2.log 2> $O/c2.err; echo "c2 rc=$?" >> $O/summary.txt
fi
cat $O/summary.txt; grep -h real $O/*.err | head; grep -c '^{' $O/*.log
