#!/bin/bash
# round 2, experiment 1: item-table schedule (CGPTB_TMA_SCHED) and per-round L2 hints (CGPTB_TMA_HINT 3/4) of k_dhop_f32_tma:
# correctness on small lattices, timing at 32^3x64x12, DRAM bytes per launch (ncu) per variant
cd "$(dirname "$0")/.."
ROUNDS=${ROUNDS:-2} STEPS=${STEPS:-300} python tools/tma_check.py > gpurun_out/r2_exp1_check.log 2>&1
grep -E "CHECK RESULT|FAIL|TIME" gpurun_out/r2_exp1_check.log
for v in "SCHED=0" "SCHED=1" "SCHED=1 HINT=3" "SCHED=1 HINT=4" "SCHED=1 GRID=144 HINT=3"; do
  envs=""
  for kv in $v; do envs="$envs CGPTB_TMA_$kv"; done
  tag=$(echo $v | tr ' =' '__')
  env $envs ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
     -k regex:k_dhop_f32_tma -s 8 -c 2 --csv --log-file gpurun_out/r2_exp1_dram_$tag.csv \
     python bench.py --steps 3 --warmup 3 --preheat 0 --no-e2e --no-cpu --no-cg --no-parity --no-kernels --no-solve > /dev/null 2>&1
  echo "== $v"; grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/r2_exp1_dram_$tag.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
