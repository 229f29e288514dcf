#!/bin/bash
# correctness + timing ablations (+ optional ncu capture with NCU=1) of the TMA sweep kernel; run under gpurun
if [ -z "$NOCHECK" ]; then
CHECK_ONLY=1 timeout 300 python tools/tma_check.py > gpurun_out/tma_check.log 2>&1; echo check rc=$?; grep -c "ok$" gpurun_out/tma_check.log; grep -v "ok$" gpurun_out/tma_check.log | tail -5
fi
export TIME_ONLY=1
if [ -n "$NCU" ]; then
ROUNDS=1 STEPS=3 VARIANTS="full:CGPTB_TMA_TRL=16" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 1 -o gpurun_out/tma_full -f python tools/tma_check.py > gpurun_out/ncu_tma.log 2>&1
echo ncu rc=$?
ROUNDS=1 STEPS=3 VARIANTS="mem:CGPTB_ABLATE=2" timeout 600 ncu --set full --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 1 -o gpurun_out/tma_mem -f python tools/tma_check.py > gpurun_out/ncu_tma_mem.log 2>&1
ROUNDS=1 STEPS=3 VARIANTS="cmp:CGPTB_ABLATE=1" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 1 -o gpurun_out/tma_cmp -f python tools/tma_check.py > gpurun_out/ncu_tma_cmp.log 2>&1
fi
if [ -z "$NOTIME" ]; then
ROUNDS=${ROUNDS:-2} STEPS=200 ABLATE=1 TRLS=${TRLS:-16,32} timeout 300 python tools/tma_check.py 2>&1 | grep TIME
fi
