#!/bin/bash
# Round 2, GPU call 1 (one B200): the whole GPU suite with lattice-relative positions and the bulk-copy staged gathers,
# the C2 / C3 full-size oracle fixtures, the new bench line, the reference arm, A/B of the kernel knobs, launch list,
# one ncu --set full capture of the new gather kernels.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_call1.sh'
set -u
mkdir -p gpurun_out
o=gpurun_out/r2a
timeout 120 python -m pytest tests/test_tma_gather.py -m gpu -q -p no:cacheprovider --timeout 100 > ${o}_tma.log 2>&1
echo "tma rc=$?" >> ${o}_tma.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 400 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > ${o}_ref.json 2> ${o}_ref_err.log
timeout 200 python tools/tune_eval.py 256 base brick_zmerge=1 gather_tma=0 gather_seg=32 gather_seg=128 base \
  brick_zmerge=1,gather_seg=32 > ${o}_tune.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-paint-bench > ${o}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_tma -c 4 -o ${o}_full_gather \
  python tools/one_eval.py 256 1 > ${o}_full_gather.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:brick_scatter -s 20 -c 2 -o ${o}_full_brick \
  python tools/one_eval.py 256 1 > ${o}_full_brick.log 2>&1
tail -3 ${o}_tma.log ${o}_pytest_gpu.log; head -c 400 ${o}_bench.json; echo; cat ${o}_tune.log
