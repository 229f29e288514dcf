#!/bin/bash
# round-end evidence, run under gpurun: GPU tests, bench line, ncu launch list of the bench command, ncu --set full of the
# dominant kernel.  Outputs land in gpurun_out/; tools/ncu_keys.py turns the .ncu-rep into the summaries under profiles/.
R=${R:-r1b}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${R}_gpu_tests.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
python bench.py --impl reference --steps 20 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cg > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_dhop_f32_tma -s 6 -c 1 -o gpurun_out/${R}_ncu_dhop_tma -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-cg > /dev/null 2>&1
cat gpurun_out/${R}_gpu_tests.log; cat gpurun_out/${R}_bench.json; cat gpurun_out/${R}_bench_reference.json; ls -la gpurun_out/${R}_*
