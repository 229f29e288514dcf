#!/bin/bash
# round-2 evidence, run under gpurun (one GPU): smoke(), both bench arms, ncu launch list of the bench command (library kernels
# only).  Outputs land in gpurun_out/; the files named in profiles/README.md are copied from there.
cd "$(dirname "$0")/.."
R=${R:-r2f}
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo smoke rc=$?; tail -3 gpurun_out/${R}_smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench.err
python bench.py > gpurun_out/${R}_bench.json 2>> gpurun_out/${R}_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
  python bench.py --steps 3 --warmup 3 --preheat 0 --no-e2e --no-cpu --no-cg --no-parity --no-kernels --no-solve --no-copy-peak > /dev/null 2>&1
grep -c k_dhop_f32_tma gpurun_out/${R}_launches.csv
python - <<PY
import json
d = json.loads(open("gpurun_out/${R}_bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d["roofline"]["frac"], d["roofline"].get("frac_first_window"), d["roofline"].get("peak_sustained_copy"), d["roofline"].get("frac_of_sustained_copy"))
print("e2e", d.get("e2e", {}).get("value"), "cpu", d.get("cpu_baseline", {}).get("value"), "cg", d.get("eo_cg", {}).get("ms_per_iteration"), "solve", (d.get("time_to_solve") or {}).get("seconds"))
PY
