#!/bin/bash
mkdir -p gpurun_out
( AMRB_VARIANT=13 timeout 600 python -m pytest tests -m gpu -x -q -k "c3_euler or r3_s16 or r3_s8_h1_d5_euler or amr3d" ) > gpurun_out/k_pytest_v13.log 2>&1
for v in 0 13; do
  echo "== variant $v"
  AMRB_VARIANT=$v bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m r3_s16_h1_euler_L5
done > gpurun_out/k_workloads.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:euler3d_march -s 5 -c 1 -f -o gpurun_out/k_march3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L5m > gpurun_out/k_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/k_launches_3d.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L5m > /dev/null 2>&1
echo done
