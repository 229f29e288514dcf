#!/bin/bash
# does the slow mode of the CG (10-14 ms per iteration instead of 6.5) come with different clocks / power?
cd "$(dirname "$0")/.."
for i in 1 2 3 4 5 6 7 8; do
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv,noheader,nounits -lms 100 > gpurun_out/${TAG}_smi_$i.csv 2>/dev/null &
  SMI=$!
  CG_MAXITER=250 python tools/cg_bench.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CG run $i', round(d['ms_per_iteration'],3))"
  kill $SMI
  python - <<PY
import statistics
rows=[l.strip().split(", ") for l in open("gpurun_out/${TAG}_smi_$i.csv") if l.strip()]
busy=[r for r in rows if float(r[2])>500]
if busy:
    print("   samples under load", len(busy), "sm MHz median", statistics.median(float(r[0]) for r in busy), "mem MHz", statistics.median(float(r[1]) for r in busy), "power W", statistics.median(float(r[2]) for r in busy), "reasons", sorted(set(r[3] for r in busy)))
PY
done
