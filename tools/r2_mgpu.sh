#!/bin/bash
# 2+ GPU check of the halo paths: bench parity + timing under torchrun with the peer-to-peer halo (default) and with NCCL
cd "$(dirname "$0")/.."
N=${N:-2}
for mode in p2p nccl; do
  echo "== halo $mode, $N ranks"
  CGPTB_HALO=$mode timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) \
    bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-cpu ${BENCH_FLAGS:---no-cg --no-kernels --no-solve} > gpurun_out/${TAG:-r2}_mgpu_${mode}_n$N.json 2> gpurun_out/${TAG:-r2}_mgpu_${mode}_n$N.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG:-r2}_mgpu_${mode}_n$N.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "parity")}, d.get("first_window", {}).get("ms_per_step"), d.get("eo_cg"))
except Exception as e:
    print("no json:", e)
PY
  tail -3 gpurun_out/${TAG:-r2}_mgpu_${mode}_n$N.err
done
