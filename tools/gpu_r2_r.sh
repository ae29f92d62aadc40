#!/bin/bash
# round-2 GPU session R (1 GPU): parity of the dense Euler body options, one CTA per SM experiment, ncu of the
# prefetch variant, fp64 pipe micro-benchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -m gpu -k "c3" > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 46"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary >> $O/dev_bench.log 2>&1
done
nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pipe_bench.cu -o /tmp/pipe_bench && /tmp/pipe_bench > $O/pipe_bench.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:euler3d_dense -s 3 -c 1 -o $O/prof_v41 \
   python bench.py --workload r3_s8_h1_euler_L6 --storage 1 --variant 41 --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_v41.log 2>&1
tail -n 4 $O/t.log; cat $O/summary.txt; cat $O/pipe_bench.txt; grep -E '^(\{|==)' $O/dev_bench.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('=='): print(l.strip(), end=' '); continue
    d=json.loads(l); print(d['config']['workload'][-24:], '%.4f ms frac %.3f'%(d['ms_per_step'], d['roofline']['frac']))"
