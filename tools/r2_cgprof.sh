#!/bin/bash
# kernel-by-kernel times of one CG iteration on the unfused (TMA stencil + sweep kernels) path, and a full capture of the sweep kernel
cd "$(dirname "$0")/.."
env CG_MAXITER=6 ${CGENV:-CGPTB_NO_EPILOGUE=1} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:^k_" -c 400 --csv --log-file gpurun_out/${TAG}_cg_launches.csv python tools/cg_bench.py > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/${TAG}_cg_launches.csv")) if len(r) > 5]
hdr = rows[0]; i_k = hdr.index("Kernel Name"); i_m = hdr.index("Metric Name"); i_v = hdr.index("Metric Value"); i_id = hdr.index("ID")
per = {}
for r in rows[1:]:
    per.setdefault(r[i_id], {"k": r[i_k]})[r[i_m]] = float(r[i_v].replace(",", ""))
ids = sorted(per, key=int)
tail = ids[-40:]
for i in tail:
    p = per[i]
    print(i, p["k"][:60], "%.1f us" % (p.get("gpu__time_duration.sum", 0) / 1e3), "%.2f GB" % ((p.get("dram__bytes_read.sum", 0) + p.get("dram__bytes_write.sum", 0)) / 1e9))
PY
#CG_MAXITER=3 CGPTB_NO_EPILOGUE=1 ncu --set full --clock-control none -k regex:k_s_sweep -s 4 -c 1 -o gpurun_out/${TAG}_ncu_sweep -f python tools/cg_bench.py > /dev/null 2>&1
