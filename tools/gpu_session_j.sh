#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/j_pytest.log 2>&1
( AMRB_QUEUE=0 timeout 600 python -m pytest tests -m gpu -x -q -k "c3_euler or r3_s16 or r3_s8_h1_d5_euler or amr3d" ) > gpurun_out/j_pytest_q0.log 2>&1
for q in 0 1; do
  echo "== queue $q"
  AMRB_QUEUE=$q bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m
done > gpurun_out/j_workloads.log 2>&1
bash tools/bench_variants.sh 0 0 >> gpurun_out/j_workloads.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:euler3d_march -s 5 -c 2 --csv --log-file gpurun_out/j_l5m.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L5m > /dev/null 2>&1
echo done
