#!/bin/bash
# round 2, experiment 2: G chunks of the fifth dimension per CTA (CGPTB_TMA_G): correctness on small lattices, timing at
# 32^3x64x12, DRAM + L2->SM bytes per launch (ncu) per variant
cd "$(dirname "$0")/.."
ROUNDS=${ROUNDS:-2} STEPS=${STEPS:-300} ABLATE=1 python tools/tma_check.py > gpurun_out/${TAG:-r2_exp2}_check.log 2>&1
grep -E "CHECK RESULT|FAIL|TIME" gpurun_out/${TAG:-r2_exp2}_check.log
for v in ${NCU_VARIANTS:-"G=1" "G=3"}; do
  envs=""
  for kv in $v; do envs="$envs CGPTB_TMA_$kv"; done
  tag=$(echo $v | tr ' =' '__')
  env $envs ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none \
     -k regex:k_dhop_f32_tma -s 8 -c 2 --csv --log-file gpurun_out/${TAG:-r2_exp2}_dram_$tag.csv \
     python bench.py --steps 3 --warmup 3 --preheat 0 --no-e2e --no-cpu --no-cg --no-parity --no-kernels --no-solve > /dev/null 2>&1
  echo "== $v"; grep -E "dram__bytes|gpu__time|hit_rate|lts__t_sectors|wavefronts|cycles_active|inst_executed" gpurun_out/${TAG:-r2_exp2}_dram_$tag.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"'
done
