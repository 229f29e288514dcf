#!/bin/bash
# 2 GPUs: CG on a decomposed lattice (sweep epilogue + fused update): iteration counts against the oracle, time per iteration
cd "$(dirname "$0")/.."
run() { timeout ${TMO:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
run tests/mgpu_check.py --mpi 1.1.1.2 --Ls 8 --only cg > gpurun_out/${TAG}_cg.log 2>&1; echo "mgpu cg rc=$?"; grep -E "cg |FAIL|PASSED" gpurun_out/${TAG}_cg.log | tail -6
run bench.py --gpus 2 --steps 100 --warmup 10 --preheat 0.5 --no-e2e --no-cpu --no-kernels --no-solve --no-parity --no-copy-peak --cg-iterations 50 2> gpurun_out/${TAG}_b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['eo_cg'])"
