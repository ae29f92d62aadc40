#!/bin/bash
# round-2 GPU session Q (1 GPU): dense Euler body options (prologue prefetch, one basic block per plane, early z)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_parity.py -q -m gpu -k "c3 or interior or dense" > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 41" "r3_s8_h1_euler_L6 1 42" "r3_s8_h1_euler_L6 1 43" "r3_s8_h1_euler_L6 1 44" "r3_s8_h1_euler_L6 1 45" "r3_s8_h1_euler_L5m 1 0" "r3_s8_h1_euler_L5m 1 41" "r3_s8_h1_euler_L5m 1 44" "r3_s16_h1_euler_L5 1 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
tail -n 4 $O/t.log; cat $O/summary.txt; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
