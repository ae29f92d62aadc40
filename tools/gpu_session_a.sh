#!/bin/bash
# One GPU-box session: tests, bench lines, ncu launch list + full capture of the dominant kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/a_bench_c2.json 2> gpurun_out/a_bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler2d_march -s 5 -c 2 -f -o gpurun_out/a_march python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_full.log 2>&1
bash tools/bench_workloads.sh r3_s8_h1_euler_L5m r3_s8_h1_euler_L6 r3_s16_h1_euler_L5 r3_s8_h1_adv_L6 r3_s16_h1_adv_L5 r2_s64_h1_adv_L5m r2_s10_h2_adv_L7 > gpurun_out/a_workloads.log 2>&1
echo done
