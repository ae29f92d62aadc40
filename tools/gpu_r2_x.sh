#!/bin/bash
# round-2 GPU session X (N GPUs): fused wait + unpack, one fence per CTA in the push: multi-GPU tests + bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2x_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
fi
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2.log 2> $O/bench_c2.err; echo "c2 rc=$?" >> $O/summary.txt
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_c3.log 2> $O/bench_c3.err; echo "c3 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 3 $O/t.log 2>/dev/null; grep -h -v "OMP_NUM\|\*\*\*\*\|^$\|unbatched" $O/*.err | tail -n 10
for f in $O/bench_c2.log $O/bench_c3.log; do grep '^{' $f | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][:12], d['n_gpus'], '%.4f ms'%d['ms_per_step'], '%.3e'%d['value'], d.get('parity',{}).get('state_max_err_over_field_max'), d.get('parity',{}).get('state_bit_identical'), d.get('ms_per_step_by_phase') or d['config'].get('ms_per_step_by_phase'))"; done
