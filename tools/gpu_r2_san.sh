#!/bin/bash
# round-2 GPU session: compute-sanitizer (memcheck, racecheck) over the new kernel families
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2san; mkdir -p $O
SEL='tests/test_gpu_parity.py::test_device_matches_reference_dump tests/test_device_regrid.py tests/test_active_amr.py'
K='interior or adv or c1_adv or wrap2d or regrid or active'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest $SEL -q -m gpu -k "$K" -x > $O/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/summary.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py::test_device_matches_reference_dump -q -m gpu -k "interior or c1_adv or wrap2d or adv3d" -x > $O/racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k "one_process and (p2p or copy) and (interior or 2d)" -x > $O/memcheck_p2p.txt 2>&1; echo "memcheck p2p rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 6 $O/memcheck.txt $O/racecheck.txt $O/memcheck_p2p.txt
