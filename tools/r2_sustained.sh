#!/bin/bash
# sustained (power-capped) A/B of kernel variants through bench.py (2 s pre-heat + 300 timed steps each), same box
cd "$(dirname "$0")/.."
for v in ${VARIANTS:-"G=1" "G=3" "G=1" "G=3"}; do
  envs=""
  for kv in $v; do envs="$envs CGPTB_TMA_$kv"; done
  env $envs python bench.py --steps 300 --no-e2e --no-cpu --no-cg --no-kernels --no-solve --no-parity > gpurun_out/${TAG}_sus.json 2> gpurun_out/${TAG}_sus.err
  python - "$v" <<PY
import json, sys
d = json.loads(open("gpurun_out/${TAG}_sus.json").read().strip().splitlines()[-1])
print(sys.argv[1], "sustained", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "first", round(d["first_window"]["ms_per_step"], 4), "clk", d["clocks"]["sm_mhz"], d["clocks"]["power_w"])
PY
done
