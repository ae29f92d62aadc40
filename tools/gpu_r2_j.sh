#!/bin/bash
# round-2 GPU session J (1 GPU): branch-free 3D advection loop
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2j; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_active_amr.py tests/test_multigpu_gpu.py -q -m gpu -x > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_adv_L6 1 0" "r3_s8_h1_adv_L6 1 31" "r3_s16_h1_adv_L5 1 0" "r3_s8_h1_adv_L5m 1 0" "r2_s32_h1_adv_L7_d9 0 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
timeout 600 python bench.py --workload c5 --steps 50 --warmup 10 > $O/bench_c5.log 2> $O/bench_c5.err; echo "c5 rc=$?" >> $O/summary.txt
tail -n 4 $O/t.log; cat $O/summary.txt; grep '^{' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
