#!/bin/bash
# round-2 GPU session Z (1 GPU): upwind form of the 3D advection kernel: parity + workloads + C5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2z; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_active_amr.py tests/test_device_regrid.py tests/test_multigpu_gpu.py -q -m gpu -k "adv or active or regrid or interior" > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_adv_L6 1 0" "r3_s16_h1_adv_L5 1 0" "r3_s8_h1_adv_L5m 1 0" "c5 1 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
cat $O/summary.txt; tail -n 3 $O/t.log; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f val %.3e'%(d['ms_per_step'], d['roofline']['frac'], d['value']))"
