#!/bin/bash
# round-2 GPU session G (1 GPU): 2D advection kernel, active AMR, N=8 weak mesh reproduction, variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2g; mkdir -p $O
timeout 600 python tools/debug_c2_w8.py 8 > $O/debug_w8.log 2>&1; echo "debug w8 rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_active_amr.py tests/test_gpu_parity.py tests/test_dropin_examples.py tests/test_dropin_drivers.py -q -m gpu > $O/t_adv.log 2>&1; echo "adv2d/amr/dropin tests rc=$?" >> $O/summary.txt
for cfg in "r2_s64_h1_adv_L6 0 0" "r2_s64_h1_adv_L6 0 10" "r2_s10_h2_adv_L9_d9 0 0" "r2_s10_h2_adv_L9_d9 0 10" "r2_s16_h1_adv_L8_d9 0 0" "r2_s32_h1_adv_L7_d9 0 0" "r3_s8_h1_euler_L6 1 25" "r3_s8_h1_euler_L6 1 26" "r3_s8_h1_euler_L6 1 27" "r3_s8_h1_euler_L6 1 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
timeout 600 python bench.py --workload c5 --steps 50 --warmup 10 > $O/bench_c5.log 2> $O/bench_c5.err; echo "c5 rc=$?" >> $O/summary.txt
tail -n 5 $O/t_adv.log; cat $O/summary.txt; cat $O/debug_w8.log
