#!/bin/bash
# round-2 GPU session AA (1 GPU): ncu of the upwind advection kernels, sanitizer over the kernels changed last
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2aa; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k regex:advect3d_dense -s 3 -c 1 -o $O/prof_advect3d_dense \
   python bench.py --workload r3_s8_h1_adv_L6 --storage 1 --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_adv3d.log 2>&1; echo "ncu adv3d rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:advect2d -s 3 -c 1 -o $O/prof_advect2d_64 \
   python bench.py --workload r2_s64_h1_adv_L6 --storage 0 --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_adv2d.log 2>&1; echo "ncu adv2d rc=$?" >> $O/summary.txt
SEL='tests/test_gpu_parity.py::test_device_matches_reference_dump tests/test_gpu_parity.py::test_device_matches_oracle_on_bench_shapes tests/test_active_amr.py'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest $SEL -q -m gpu -k "adv or active" -x > $O/memcheck.txt 2>&1; echo "memcheck rc=$?" >> $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 5 python -m pytest tests/test_gpu_parity.py::test_device_matches_reference_dump -q -m gpu -k "adv" -x > $O/racecheck.txt 2>&1; echo "racecheck rc=$?" >> $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k "one_process" -x > $O/memcheck_cluster.txt 2>&1; echo "memcheck cluster rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 3 $O/memcheck.txt $O/racecheck.txt $O/memcheck_cluster.txt
