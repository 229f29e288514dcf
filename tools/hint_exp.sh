#!/bin/bash
CHECK_ONLY=1 timeout 300 python tools/tma_check.py > gpurun_out/tma_check.log 2>&1; echo check rc=$?; grep -c "ok$" gpurun_out/tma_check.log; grep -v "ok$" gpurun_out/tma_check.log | tail -5
export TIME_ONLY=1
V=""
for h in 0 1 2; do for t in 16 32; do V="$V;full h$h t$t:CGPTB_TMA_HINT=$h,CGPTB_TMA_TRL=$t"; done; V="$V;mem h$h:CGPTB_TMA_HINT=$h,CGPTB_ABLATE=2"; done
ROUNDS=2 STEPS=200 VARIANTS="${V#;}" timeout 300 python tools/tma_check.py 2>&1 | grep TIME
for h in 0 1 2; do
ROUNDS=1 STEPS=3 VARIANTS="full:CGPTB_TMA_HINT=$h" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 2 python tools/tma_check.py 2>&1 | grep -E "dram__|gpu__time|hit_rate" | sed "s/^/hint=$h /"
done
