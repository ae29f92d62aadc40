#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1
bash tools/bench_workloads.sh r3_s8_h1_adv_L6 r3_s16_h1_adv_L5 r2_s64_h1_adv_L5m r2_s10_h2_adv_L7 > gpurun_out/m_workloads.log 2>&1
for d in ref_bench_fvm_solver_integration ref_bench_fvm_solver_integration3D ref_bench_fvm_solver_integration_active_amr; do
  echo "== $d"; ( time timeout 600 examples/_build/$d ) 2>&1 | grep -v "^Step" | tail -25
done > gpurun_out/m_dropin_drivers.log 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/m_bench_c2.json 2> gpurun_out/m_bench_c2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler2d_march -s 5 -c 2 -f -o gpurun_out/m_march python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/m_ncu_full.log 2>&1
echo done
