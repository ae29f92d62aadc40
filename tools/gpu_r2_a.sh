#!/bin/bash
# round-2 GPU session A: smoke, new parity tests, dense-vs-padded kernel timings, reference CUDA timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpu.txt 2>&1
free -g > $O/mem.txt 2>&1; nproc >> $O/mem.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "interior or oracle_on_bench" -x > $O/t_dense.log 2>&1; echo "dense tests rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_parity_full.py -q -m gpu > $O/t_full.log 2>&1; echo "full-size tests rc=$?" >> $O/summary.txt
for cfg in "r3_s8_h1_euler_L6 0 0" "r3_s8_h1_euler_L6 1 0" "r3_s8_h1_euler_L6 1 21" "r3_s8_h1_euler_L6 1 22" "r3_s8_h1_euler_L5m 1 0" "r3_s8_h1_adv_L6 0 0" "r3_s8_h1_adv_L6 1 0" "r3_s8_h1_adv_L6 1 31" "r3_s16_h1_euler_L5 0 0" "r3_s16_h1_euler_L5 1 0" "r3_s16_h1_adv_L5 0 0" "r3_s16_h1_adv_L5 1 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
echo "dev bench done" >> $O/summary.txt
timeout 120 oracle/_ref/ref_cuda_bench_2d <(python -c "
import importlib,sys; sys.path.insert(0,'.')
wl=importlib.import_module('gpu-amr_b200.workloads'); print(wl.c2_script()+'\nI\nX\nT 3\nT 20')") /tmp/o2.bin 4096 > $O/refcuda_2d.log 2>&1; echo "refcuda2d rc=$?" >> $O/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_c3.log 2> $O/bench_c3.err; echo "bench c3 rc=$?" >> $O/summary.txt
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_parity_full.py > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" >> $O/summary.txt
tail -3 $O/t_dense.log $O/t_full.log $O/t_all.log
cat $O/summary.txt
