#!/bin/bash
# ncu --set full captures of the TMA sweep kernel for the G variants named in $GS (default "1 3"); run under gpurun
cd "$(dirname "$0")/.."
export TIME_ONLY=1
for G in ${GS:-1 3}; do
ROUNDS=1 STEPS=3 VARIANTS="full:CGPTB_TMA_G=$G" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_dhop_f32_tma -s 4 -c 1 -o gpurun_out/${TAG:-r2}_ncu_G$G -f python tools/tma_check.py > gpurun_out/${TAG:-r2}_ncu_G$G.log 2>&1
echo "G=$G ncu rc=$?"
done
