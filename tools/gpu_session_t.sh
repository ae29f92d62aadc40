#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "adv" ) > gpurun_out/t_pytest.log 2>&1
bash tools/bench_workloads.sh r3_s8_h1_adv_L6 r3_s16_h1_adv_L5 > gpurun_out/t_workloads.log 2>&1
echo done
