#!/bin/bash
# 8-GPU evidence: bench line T-split and T x Z, configs[3] (clover solve 48^3x96 on 1.1.2.4), configs[4] (Moebius 64^3x128 12 columns)
cd "$(dirname "$0")/.."
N=${N:-8}
run() { timeout ${TMO:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) "$@"; }
show() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "n_gpus", "true_residual", "iterations_last_column", "inner_iterations_last_cycle", "converged")}
    keep["parity"] = (d.get("parity") or {}).get("rel_err")
    keep["first"] = (d.get("first_window") or {}).get("ms_per_step")
    keep["cg"] = (d.get("eo_cg") or {}).get("ms_per_iteration")
    keep["clk"] = (d.get("clocks") or {}).get("sm_mhz")
    print(sys.argv[1], keep)
except Exception as e:
    print(sys.argv[1], "no json:", e)
PY
}
run bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-cpu --no-kernels --no-solve --cg-iterations 50 > gpurun_out/${TAG}_t.json 2> gpurun_out/${TAG}_t.err; show gpurun_out/${TAG}_t.json
[ -n "$SKIP_TZ" ] || run bench.py --gpus $N --mpi 1.1.2.$((N/2)) --steps 300 --warmup 10 --no-e2e --no-cpu --no-kernels --no-solve --cg-iterations 50 > gpurun_out/${TAG}_tz.json 2> gpurun_out/${TAG}_tz.err; [ -n "$SKIP_TZ" ] || show gpurun_out/${TAG}_tz.json
run bench.py --gpus $N --config clover_solve --mpi 1.1.2.$((N/2)) ${CLOVER_GRID:+--grid $CLOVER_GRID} > gpurun_out/${TAG}_c3.json 2> gpurun_out/${TAG}_c3.err; show gpurun_out/${TAG}_c3.json
run bench.py --gpus $N --config mobius_prop ${MOBIUS_GRID:+--grid $MOBIUS_GRID} > gpurun_out/${TAG}_c4.json 2> gpurun_out/${TAG}_c4.err; show gpurun_out/${TAG}_c4.json
grep -h -i "error\|Traceback" gpurun_out/${TAG}_*.err | head -5
