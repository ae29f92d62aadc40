#!/bin/bash
# round-2 GPU session T (1 GPU): wave-form Rusanov flux -- whole GPU suite + Euler workloads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2t; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 28" "r3_s8_h1_euler_L6 1 41" "r3_s8_h1_euler_L5m 1 0" "r3_s16_h1_euler_L5 1 0" "r3_s8_h1_euler_L6 0 0" "c2 0 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
tail -n 4 $O/t.log; cat $O/summary.txt; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
