#!/bin/bash
# Round 2, GPU call 26: the two fusions of SURVEY 8f rows 1 and 3 timed against the compositions they replace.
set -u
mkdir -p gpurun_out
timeout 400 python tools/fused_probe.py 256 > gpurun_out/r2z_fused.log 2>&1
cat gpurun_out/r2z_fused.log | tail -8
