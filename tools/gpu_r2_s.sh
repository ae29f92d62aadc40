#!/bin/bash
# round-2 GPU session S (1 GPU): single-load table pieces, tickets one task ahead; whole GPU suite, variants, sanitizer
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2s; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -x > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 41" "r3_s8_h1_euler_L6 1 44" "r3_s8_h1_euler_L5m 1 0" "r3_s8_h1_euler_L5m 1 41" "r3_s16_h1_euler_L5 1 0" \
           "r3_s8_h1_adv_L6 1 0" "r3_s16_h1_adv_L5 1 0" "r3_s8_h1_adv_L5m 1 0" "r3_s8_h1_euler_L6 0 0" "c2 0 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
SEL='tests/test_gpu_parity.py::test_device_matches_reference_dump tests/test_active_amr.py'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest $SEL -q -m gpu -k "interior or adv or active" -x > $O/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/summary.txt
tail -n 4 $O/t.log; cat $O/summary.txt; tail -n 3 $O/memcheck.txt; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
