#!/bin/bash
# round-2 final check on one GPU: whole GPU suite, the driver's two bench arms and smoke(), exactly as the driver runs them
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2final; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/ref.log 2> $O/ref.err; echo "ref rc=$?" >> $O/summary.txt
( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/ours.log 2> $O/ours.err; echo "ours rc=$?" >> $O/summary.txt
( time python3 -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 3 $O/pytest_gpu.log; grep -h real $O/*.err; tail -n 5 $O/smoke.log
