#!/bin/bash
mkdir -p gpurun_out
for r in 0 1 3 0 1 3; do echo -n "ring $r "; AMRB_RING=$r bash tools/bench_variants.sh 3; done > gpurun_out/h_ring.log 2>&1
( AMRB_RING=1 timeout 600 python -m pytest tests -m gpu -x -q -k "c2 or r2_s64" ) > gpurun_out/h_pytest_ring1.log 2>&1
echo done
