#!/bin/bash
# Round 2, GPU call 21: where the slab-decomposed model loses time against the single-GPU engine at ONE rank (256^3):
# its launch list, and its timing with the default and a 12-plane halo.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2u
timeout 200 python tools/slab_bench.py --mesh 256 --model --steps 3 --warmup 2 > ${o}_slab1.json 2> ${o}_slab1_err.log
timeout 200 python tools/slab_bench.py --mesh 256 --model --steps 3 --warmup 2 --halo 12 > ${o}_slab1_h12.json 2> ${o}_slab1_h12_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file ${o}_launches.csv \
  python tools/slab_bench.py --mesh 256 --model --steps 1 --warmup 1 > ${o}_ncu.log 2>&1
grep "^{" ${o}_slab1.json | head -c 600; echo; grep "^{" ${o}_slab1_h12.json | head -c 600; echo; tail -n 2 ${o}_ncu.log
