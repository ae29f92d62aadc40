#!/bin/bash
# round-2 GPU session H (2 GPUs): shim binding, 2D advection pairs, tolerance parity, active AMR over NCCL/p2p
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
O=gpurun_out/r2h_n$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655"
timeout 900 python -m pytest tests/test_shim_binding.py tests/test_gpu_parity.py tests/test_dropin_examples.py -q -m gpu > $O/t_shim_adv.log 2>&1; echo "shim+adv tests rc=$?" >> $O/summary.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -k "active_amr" > $O/t_amr_multi.log 2>&1; echo "active amr multi rc=$?" >> $O/summary.txt
for cfg in "r2_s64_h1_adv_L6 0 0" "r2_s64_h1_adv_L6 0 41" "r2_s64_h1_adv_L6 0 42" "r2_s32_h1_adv_L7_d9 0 41" "r2_s32_h1_adv_L7_d9 0 42" "r2_s32_h1_adv_L7_d9 0 0" "r2_s16_h1_adv_L8_d9 0 0" "r2_s10_h2_adv_L9_d9 0 0" "r2_s8_h1_adv_L9_d9 0 0"; do
  set -- $cfg
  echo "== $cfg" >> $O/dev_bench.log
  timeout 300 python bench.py --workload $1 --storage $2 --variant $3 --steps 10 --warmup 3 --no-cpu-baseline >> $O/dev_bench.log 2>&1
done
timeout 600 $TR bench.py --gpus $N --workload c2 --steps 20 --warmup 5 > $O/bench_c2.log 2> $O/bench_c2.err; echo "c2 rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c3L4 --steps 20 --warmup 5 > $O/bench_c3L4.log 2> $O/bench_c3L4.err; echo "c3L4 rc=$?" >> $O/summary.txt
timeout 600 $TR bench.py --gpus $N --workload c5 --steps 50 --warmup 10 > $O/bench_c5.log 2> $O/bench_c5.err; echo "c5 rc=$?" >> $O/summary.txt
tail -n 5 $O/t_shim_adv.log $O/t_amr_multi.log; cat $O/summary.txt; grep -h -v "OMP_NUM\|\*\*\*\*\|^$" $O/*.err | tail -n 20
