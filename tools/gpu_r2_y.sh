#!/bin/bash
# round-2 GPU session Y (1 GPU): per-entry layer counts + fused receive kernel through the one-process cluster tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2y; mkdir -p $O
timeout 1200 python -m pytest tests/test_multigpu_gpu.py tests/test_active_amr.py -q -m gpu > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k "one_process and p2p and (interior or 2d)" -x > $O/memcheck_p2p.txt 2>&1; echo "memcheck p2p rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 4 $O/t.log; tail -n 3 $O/memcheck_p2p.txt
