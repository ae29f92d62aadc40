#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/l_pytest.log 2>&1
{ bash tools/bench_variants.sh 0 0; bash tools/bench_workloads.sh r3_s8_h1_euler_L6 r3_s8_h1_euler_L5m; } > gpurun_out/l_workloads.log 2>&1
echo done
