#!/bin/bash
# Round-end check on one B200: GPU tests, smoke, both bench arms; optional compute-sanitizer pass (SAN=1)
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/z_pytest_gpu.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err
( time timeout 300 examples/_build/ref_bench_fvm_solver_integration_active_amr ) 2>&1 | grep -v "^Step" | tail -22 > gpurun_out/z_active_amr.log
if [ "$SAN" = "1" ]; then
K='(test_device_matches_reference_dump and fused and not fused_v1 and (c3_euler or amr3d_euler)) or (test_device_matches_oracle_on_bench_shapes and (r2_s64 or r3_s8_h1_d5_euler))'
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" ) > gpurun_out/z_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/z_sanitizer_memcheck.log
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" ) > gpurun_out/z_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/z_sanitizer_racecheck.log
fi
echo done
