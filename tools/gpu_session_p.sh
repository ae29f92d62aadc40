#!/bin/bash
mkdir -p gpurun_out
( AMRB_VARIANT=11 timeout 600 python -m pytest tests -m gpu -x -q -k "r3_s16 or c3 or r3_s8_h1_d5_euler" ) > gpurun_out/p_pytest_v11.log 2>&1
for v in 0 11; do
  echo "== variant $v"
  AMRB_VARIANT=$v bash tools/bench_workloads.sh r3_s16_h1_euler_L5 r3_s16_h1_euler_L4m
done > gpurun_out/p_workloads.log 2>&1
echo done
