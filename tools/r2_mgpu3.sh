#!/bin/bash
cd "$(dirname "$0")/.."
N=${N:-2}
run() { timeout ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
for pack in copy direct; do
  echo "== p2p pack=$pack"
  CGPTB_HALO_PACK=$pack CGPTB_HALO_TIMING=1 run bench.py --gpus $N --steps 100 --warmup 10 --preheat 0.5 --no-e2e --no-cpu --no-cg --no-kernels --no-solve --no-parity > /dev/null 2> gpurun_out/${TAG}_timing_$pack.err
  grep "halo timing" gpurun_out/${TAG}_timing_$pack.err | tail -2
  CGPTB_HALO_PACK=$pack run bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-cpu --no-cg --no-kernels --no-solve > gpurun_out/${TAG}_bench_$pack.json 2> gpurun_out/${TAG}_bench_$pack.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_$pack.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d["parity"]["rel_err"], d.get("first_window", {}).get("ms_per_step"))
PY
done
