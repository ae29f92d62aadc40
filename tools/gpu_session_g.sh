#!/bin/bash
mkdir -p gpurun_out
for w in r3_s8_h1_euler_L4 r3_s8_h1_euler_L5; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_op_read_hit_rate.pct,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum --clock-control none -k regex:euler3d_march -s 5 -c 2 --csv --log-file gpurun_out/g_$w.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/g_$w.log 2>&1
done
echo done
