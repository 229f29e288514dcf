#!/bin/bash
# 2-GPU: (a) phase timing of the decomposed Dslash, (b) clover + mixed-precision solve parity on T- and Z-split lattices
cd "$(dirname "$0")/.."
run() { timeout ${TMO:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
CGPTB_HALO_TIMING=1 run bench.py --gpus 2 --steps 100 --warmup 10 --preheat 0.5 --no-e2e --no-cpu --no-cg --no-kernels --no-solve --no-parity > gpurun_out/${TAG}_timing.json 2> gpurun_out/${TAG}_timing.err
grep "halo timing" gpurun_out/${TAG}_timing.err | tail -4
for mpi in 1.1.1.2 1.1.2.1; do
  run tests/mgpu_check.py --mpi $mpi --Ls 8 --only clover,solve > gpurun_out/${TAG}_clover_$mpi.log 2>&1
  echo "mpi $mpi rc=$?"; grep -E "FAIL|PASSED|solve|Error|error" gpurun_out/${TAG}_clover_$mpi.log | tail -12
done
