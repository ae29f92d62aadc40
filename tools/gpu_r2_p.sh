#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2p; mkdir -p $O
for v in 0 26; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:euler3d_dense -s 4 -c 1 -o $O/prof_v$v \
   python bench.py --workload r3_s8_h1_euler_L6 --storage 1 --variant $v --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_v$v.log 2>&1
done
ls -la $O
