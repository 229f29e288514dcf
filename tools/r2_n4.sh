#!/bin/bash
cd "$(dirname "$0")/.."
N=4
run() { timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
for mpi in 1.1.1.4 1.1.2.2; do
run bench.py --gpus $N --mpi $mpi --steps 300 --warmup 10 --no-e2e --no-cpu --no-kernels --no-solve --no-copy-peak --cg-iterations 50 > gpurun_out/${TAG}_$mpi.json 2> gpurun_out/${TAG}_$mpi.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_$mpi.json").read().strip().splitlines()[-1])
    print("$mpi", d["ms_per_step"], d["first_window"]["ms_per_step"], d["parity"]["rel_err"], d["eo_cg"]["ms_per_iteration"], d["eo_cg"]["ms_per_iteration_all_tries"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$mpi no json", e)
PY
tail -2 gpurun_out/${TAG}_$mpi.err | cut -c1-200
done
