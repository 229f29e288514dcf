#!/bin/bash
export TIME_ONLY=1
for mb in 32 64 96; do
export CGPTB_L2_PERSIST_MB=$mb
for v in "hint1:CGPTB_TMA_HINT=1,CGPTB_L2_PERSIST_MB=$mb" ; do
ROUNDS=1 STEPS=3 VARIANTS="$v" timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 1 python tools/tma_check.py 2>&1 | grep -E "dram__|gpu__time|lts__|persisting" | awk -v n="$v" '{printf "%s %s %s | ", n, $1, $3} END {print ""}'
done
ROUNDS=2 STEPS=200 VARIANTS="full h0:CGPTB_TMA_HINT=0,CGPTB_L2_PERSIST_MB=$mb;full h1:CGPTB_TMA_HINT=1,CGPTB_L2_PERSIST_MB=$mb" timeout 300 python tools/tma_check.py 2>&1 | grep -E "TIME|persisting"
done
