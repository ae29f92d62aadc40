#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/r_pytest_gpu.log 2>&1
( AMRB_DEVICE_TOPOLOGY=0 timeout 600 python -m pytest tests -m gpu -x -q -k "reference_dump and fused and not fused_v1" ) > gpurun_out/r_pytest_hosttopo.log 2>&1
for d in ref_bench_fvm_solver_integration_active_amr; do
  echo "== $d"; ( time timeout 600 examples/_build/$d ) 2>&1 | grep -v "^Step" | tail -12
done > gpurun_out/r_dropin_active_amr.log 2>&1
echo done
