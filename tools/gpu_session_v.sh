#!/bin/bash
# per-kernel achieved DRAM bandwidth of every kernel of the library (ncu, 3 metrics, no replay cost to speak of)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --clock-control none -s 200 -c 1200 --csv --log-file gpurun_out/v_kernels_active_amr.csv examples/_build/ref_bench_fvm_solver_integration_active_amr > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 4 -c 24 --csv --log-file gpurun_out/v_kernels_adv2d.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r2_s64_h1_adv_L5m > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 4 -c 24 --csv --log-file gpurun_out/v_kernels_adv3d.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_adv_L6 > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -s 4 -c 24 --csv --log-file gpurun_out/v_kernels_euler3d.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload r3_s8_h1_euler_L6 > /dev/null 2>&1
echo done
