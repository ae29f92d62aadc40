#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over parity tests that drive every fused kernel family
mkdir -p gpurun_out
K='(test_device_matches_reference_dump and fused and not fused_v1 and (c3_euler or c3_amr or amr3d_euler or adv3d or c1_adv)) or (test_device_matches_oracle_on_bench_shapes and (r2_s64 or r3_s8_h1_d5_euler or r3_s16)) or test_device_tables_equal_host_tables'
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_device_topology.py -x -q -k "$K" ) > gpurun_out/z_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/z_sanitizer_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_device_topology.py -x -q -k "$K" ) > gpurun_out/z_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/z_sanitizer_racecheck.log
echo done
