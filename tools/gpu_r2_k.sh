#!/bin/bash
# round-2 GPU session K (1 GPU): reconstruct_tree on the device
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2k; mkdir -p $O
timeout 900 python -m pytest tests/test_device_regrid.py tests/test_active_amr.py tests/test_dropin_drivers.py tests/test_dropin_examples.py -q -m gpu > $O/t.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
timeout 600 python bench.py --workload c5 --steps 50 --warmup 10 > $O/bench_c5.log 2> $O/bench_c5.err; echo "c5 rc=$?" >> $O/summary.txt
( time examples/_build/ref_bench_fvm_solver_integration_active_amr ) > $O/dropin_active_amr.txt 2>&1
tail -n 30 $O/t.log; cat $O/summary.txt; grep '^{' $O/bench_c5.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('c5 %.3e upd/s %.4f ms/step'%(d['value'], d['ms_per_step']), d['config']['patches_end'], d['config']['topology_changing_reconstructs'])"; tail -n 22 $O/dropin_active_amr.txt
