#!/bin/bash
# quick validation of the bench --config workloads on small lattices (2 GPUs)
cd "$(dirname "$0")/.."
run() { timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
python bench.py --config wilson_clover_16 --steps 200 2> gpurun_out/${TAG}_c0.err | tee gpurun_out/${TAG}_c0.json | cut -c1-600
run bench.py --gpus 2 --config clover_solve --grid 16.16.16.32 --mpi 1.1.2.1 2> gpurun_out/${TAG}_c3.err | tee gpurun_out/${TAG}_c3.json | cut -c1-700
tail -2 gpurun_out/${TAG}_c3.err
run bench.py --gpus 2 --config mobius_prop --grid 16.16.16.32 2> gpurun_out/${TAG}_c4.err | tee gpurun_out/${TAG}_c4.json | cut -c1-700
tail -2 gpurun_out/${TAG}_c4.err
# T x Z parity of the bench line's own check, 2 ranks split in Z
run bench.py --gpus 2 --mpi 1.1.2.1 --steps 100 --no-e2e --no-cpu --no-kernels --no-solve --cg-iterations 30 2> gpurun_out/${TAG}_z.err | tee gpurun_out/${TAG}_z.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step','parity','eo_cg')})"
tail -2 gpurun_out/${TAG}_z.err
