#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2z3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "adv" > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r2_s64_h1_adv_L6 0 0" "r2_s32_h1_adv_L7_d9 0 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
cat $O/summary.txt; tail -n 3 $O/t.log; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f val %.3e'%(d['ms_per_step'], d['roofline']['frac'], d['value']))"
