#!/bin/bash
# where does the energy go: sustained power and time of the compute-only and memory-only ablations against the full kernel
cd "$(dirname "$0")/.."
for v in "" "CGPTB_ABLATE=1" "CGPTB_ABLATE=2" "CGPTB_TMA_G=3" "CGPTB_ABLATE=1 CGPTB_TMA_G=3" "CGPTB_ABLATE=2 CGPTB_TMA_G=3" "CGPTB_NO_TMA=1"; do
  env $v python bench.py --steps 300 --no-e2e --no-cpu --no-cg --no-kernels --no-solve --no-parity > gpurun_out/${TAG}_p.json 2> gpurun_out/${TAG}_p.err
  python - "$v" <<PY
import json, sys
d = json.loads(open("gpurun_out/${TAG}_p.json").read().strip().splitlines()[-1])
c = d["clocks"]
print("%-32s sustained %.4f ms  clk %s MHz  power %s W  energy/step %.3f J  reasons %s" % (sys.argv[1] or "full G=1", d["ms_per_step"], c["sm_mhz"], c["power_w"], (c["power_w"] or 0) * d["ms_per_step"] * 1e-3, c["reasons"]))
PY
done
nvidia-smi --query-gpu=power.draw,clocks.sm --format=csv,noheader
