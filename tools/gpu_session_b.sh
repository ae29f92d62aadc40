#!/bin/bash
# GPU session B: parity of the 3D plane-marching kernel, its variants, workloads, ncu capture
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_pytest_gpu.log 2>&1
for v in 11 12 10; do
  ( AMRB_VARIANT=$v timeout 600 python -m pytest tests -m gpu -x -q -k "c3_euler or r3_s16 or r3_s8_h1_d5_euler or 3d" ) > gpurun_out/b_pytest_v$v.log 2>&1
done
for v in 0 11 12 10; do
  echo "== variant $v"; AMRB_VARIANT=$v bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m r3_s16_h1_euler_L5
done > gpurun_out/b_workloads.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:euler3d_march -s 5 -c 1 -f -o gpurun_out/b_march3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L5m > gpurun_out/b_ncu_full.log 2>&1
timeout 300 ./tools/tma_bench > gpurun_out/b_tma_bench.log 2>&1
echo done
