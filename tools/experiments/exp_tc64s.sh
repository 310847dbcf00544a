#!/bin/bash
# per-variant timing of the stacked halo-tile kernel (256 = one-class kernel; bits of the stacked kernel: 2 no epilogue [forward only: the data-gradient line deadlocks], 4 no wrap MMAs, 8 12 of 72 MMAs, 16 no input TMA)
for v in "$@"; do echo "== variant $v"; timeout 120 python tools/experiments/gpu_exp_tc.py v2perf $v 2>&1 | grep "fprop\|dgrad"; done
