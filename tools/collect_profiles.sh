#!/bin/bash
# round-end evidence, run under gpurun: smoke(), GPU tests, bench line, ncu launch list of the bench command (library
# kernels only), ncu --set full of the dominant kernel.  Outputs land in gpurun_out/; the summaries under profiles/ are
# made from them here (tools/ncu_keys.py prints the key metrics of a .ncu-rep).
R=${R:-r1b}
python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/${R}_smoke.log 2>&1; echo smoke rc=$?
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.log
if [ -z "$QUICK" ]; then
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
ncu --set full --import-source on --clock-control none -k regex:k_dhop_f32_tma -s 6 -c 1 -o gpurun_out/${R}_ncu_dhop_tma -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cg > /dev/null 2>&1
fi
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^k_" -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cg > /dev/null 2>&1
tail -4 gpurun_out/${R}_smoke.log; cat gpurun_out/${R}_gpu_tests.log; [ -z "$QUICK" ] && cat gpurun_out/${R}_bench.json
