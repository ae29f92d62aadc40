#!/bin/bash
# round-2 GPU session V (1 GPU): whole suite, driver rehearsal (reference arm, bench, smoke), launch list of the bench
# command, ncu --set full of the bench kernel (uniform level 6) and its DRAM bytes on the C3 bench mesh itself
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/r2v; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "tests rc=$?" >> $O/summary.txt
( time python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $O/ref.log 2> $O/ref.err; echo "ref rc=$?" >> $O/summary.txt
( time python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > $O/ours.log 2> $O/ours.err; echo "ours rc=$?" >> $O/summary.txt
( time python3 -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:amrb -c 600 --csv --log-file $O/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "launch list rc=$?" >> $O/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:euler3d_dense -s 3 -c 1 -o $O/prof_euler3d_dense \
   python bench.py --workload r3_s8_h1_euler_L6 --storage 1 --steps 4 --warmup 3 --no-cpu-baseline --no-secondary > $O/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/summary.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:euler3d_dense -s 2 -c 1 --csv --log-file $O/c3_dram.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $O/ncu_c3.log 2>&1; echo "ncu c3 rc=$?" >> $O/summary.txt
cat $O/summary.txt; tail -n 3 $O/pytest_gpu.log; grep -h real $O/*.err; tail -n 2 $O/smoke.log; tail -n 4 $O/c3_dram.csv
