#!/bin/bash
# round-2 GPU session I (1 GPU): table-prefetching advection kernels, conflict-free boundary-flux parking
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2i; mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_adv_L6 1 0" "r3_s16_h1_adv_L5 1 0" "r3_s8_h1_adv_L5m 1 0" "r2_s64_h1_adv_L6 0 0" "r2_s64_h1_adv_L6 0 41" "r2_s64_h1_adv_L6 0 42" "r2_s32_h1_adv_L7_d9 0 42" "r2_s32_h1_adv_L7_d9 0 0" "r2_s16_h1_adv_L8_d9 0 0" "r2_s10_h2_adv_L9_d9 0 0" "r2_s8_h1_adv_L9_d9 0 0" "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L5m 1 0" "r3_s16_h1_euler_L5 1 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
timeout 600 python bench.py --workload c5 --steps 50 --warmup 10 > $O/bench_c5.log 2> $O/bench_c5.err; echo "c5 rc=$?" >> $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect3d_dense -s 4 -c 1 -o $O/prof_advect3d_dense \
   python bench.py --workload r3_s8_h1_adv_L6 --storage 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full_adv.log 2>&1
tail -n 5 $O/t_all.log; cat $O/summary.txt
