#!/bin/bash
mkdir -p gpurun_out
for v in 14 15; do ( AMRB_VARIANT=$v timeout 300 python -m pytest tests -m gpu -x -q -k "c3 or r3_s8_h1_d5_euler" ) > gpurun_out/q_pytest_v$v.log 2>&1; done
for v in 0 14 15; do
  echo "== variant $v"
  AMRB_VARIANT=$v bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m
done > gpurun_out/q_workloads.log 2>&1
echo done
