#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r02_determinism5.log; : > $O
run() { env "$@" timeout 240 python tools/determinism.py C4r 120 2>&1 | grep -v "^$" | tail -8 >> $O; tail -1 $O | cut -c1-330; }
run A=1; run CB200_NO_TMA=1; run CB200_TMA_TILE=128; run CB200_TMA_FENCE=1; run CB200_TMA_PAD_KB=100
