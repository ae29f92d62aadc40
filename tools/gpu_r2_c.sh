#!/bin/bash
# round-2 GPU session C (1 GPU): ncu captures of the dense 3D kernels + launch lists (ours vs the reference CUDA
# backend), drop-in API test, full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_dropin_api.py -q -m gpu > $O/t_dropin_api.log 2>&1; echo "dropin api rc=$?" >> $O/summary.txt
# launch list of our C3-family step (dev mesh L5m: 3 levels, 6.3e4 patches) and ncu --set full of the step kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_ours_3d.csv \
   python bench.py --workload r3_s8_h1_euler_L5m --storage 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_ours_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler3d_dense -s 4 -c 1 -o $O/prof_euler3d_dense \
   python bench.py --workload r3_s8_h1_euler_L6 --storage 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full_3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:advect3d_dense -s 4 -c 1 -o $O/prof_advect3d_dense \
   python bench.py --workload r3_s8_h1_adv_L6 --storage 1 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full_adv.log 2>&1
# the reference's own CUDA kernels: launch lists on the 3D sample and on C2
python - > $O/scripts.txt <<'PY'
import importlib,sys; sys.path.insert(0,'.')
wl=importlib.import_module('gpu-amr_b200.workloads')
open('/tmp/ref3d.txt','w').write(wl.c3_script(4)+'\nI\nX\nT 2\nT 3\n')
open('/tmp/ref2d.txt','w').write(wl.c2_script()+'\nI\nX\nT 2\nT 3\n')
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_refcuda_3d.csv oracle/_ref/ref_cuda_bench_3d /tmp/ref3d.txt /tmp/o3.bin 100000 > $O/refcuda_3d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_refcuda_2d.csv oracle/_ref/ref_cuda_bench_2d /tmp/ref2d.txt /tmp/o2.bin 4096 > $O/refcuda_2d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_ours_c2.csv \
   python bench.py --workload c2 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_ours_c2.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/summary.txt
tail -n 3 $O/t_dropin_api.log $O/t_all.log; cat $O/summary.txt; ls -la $O
