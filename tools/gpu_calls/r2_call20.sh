#!/bin/bash
# Round 2, GPU call 20: ncu --set full of the two top kernels inside an evaluation; TSC / PCS evaluations timed.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2t
timeout 300 python tools/order_probe.py 256 > ${o}_orders.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"brick_stream_kernel|gather_tma_kernel" -s 44 -c 3 -o ${o}_full_top \
  python tools/one_eval.py 256 2 > ${o}_full_top.log 2>&1
cat ${o}_orders.log; tail -n 4 ${o}_full_top.log
