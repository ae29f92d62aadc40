#!/bin/bash
# round-2 GPU session D (1 GPU): peer-memory exchange in one process, drop-in API test, dense Euler occupancy variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_dropin_api.py -q -m gpu > $O/t_p2p_dropin.log 2>&1; echo "p2p+dropin tests rc=$?" >> $O/summary.txt
timeout 600 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -k c3 > $O/t_c3.log 2>&1; echo "c3 parity rc=$?" >> $O/summary.txt
for v in 0 23 24 22; do
  echo "== r3_s8_h1_euler_L6 1 $v" >> $O/dev_bench.log
  timeout 300 python bench.py --workload r3_s8_h1_euler_L6 --storage 1 --variant $v --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
tail -n 4 $O/t_p2p_dropin.log $O/t_c3.log; cat $O/summary.txt
