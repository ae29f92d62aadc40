#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/i_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/i_bench_c2.json 2> gpurun_out/i_bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/i_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler2d_march -s 5 -c 2 -f -o gpurun_out/i_march python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/i_ncu_full.log 2>&1
echo done
