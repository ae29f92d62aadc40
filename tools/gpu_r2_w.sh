#!/bin/bash
# round-2 GPU session W (1 GPU): shared-memory data pipe experiments on the dense Euler kernel; launch list of bench.py
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2w; mkdir -p $O
for cfg in "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 47" "r3_s8_h1_euler_L6 1 48" "r3_s8_h1_euler_L6 1 28" "r3_s8_h1_euler_L5m 1 0" "r3_s8_h1_euler_L5m 1 48"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
timeout 300 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -k "c3" > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:compute_dt|euler|halo_kernel|init_scalars|topology|plan_kernel|interior_copy|advect|face_|dense_|flags" -c 600 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "launch list rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 2 $O/t.log; wc -l $O/launches_bench.csv; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
