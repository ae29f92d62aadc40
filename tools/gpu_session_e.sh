#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "not dropin" ) > gpurun_out/e_pytest.log 2>&1
for tm in 0 1; do
  echo "== taskmap $tm"
  AMRB_TASKMAP=$tm bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m
  AMRB_TASKMAP=$tm bash tools/bench_variants.sh 0 3 2
done > gpurun_out/e_workloads.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:euler3d_march -s 5 -c 1 -f -o gpurun_out/e_march3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L5m > gpurun_out/e_ncu_full.log 2>&1
echo done
