#!/bin/bash
# Wilson-clover Dhop (Ls = 1): one warp per hop direction (CGPTB_DHOP_DIR=1) against the one-thread-per-site kernels (=0)
cd "$(dirname "$0")/.."
for grid in 16.16.16.16 32.32.32.64; do for prec in "" "--single"; do for dir in 0 1; do
  CGPTB_DHOP_DIR=$dir python bench.py --config wilson_clover_16 --grid $grid $prec --steps 100 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('grid $grid prec ${prec:-double} dir $dir: flushed %.4f ms  frac %.3f  resident %.4f ms  %.0f GFlop/s' % (d['ms_per_step'], d['roofline']['frac'], d['ms_per_step_without_l2_flush'], d['value']))"
done; done; done
python -m pytest tests -m gpu -x -q -k "wilson or clover or twisted or open_bc" 2>&1 | tail -3
