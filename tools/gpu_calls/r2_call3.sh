#!/bin/bash
# Round 2, GPU call 3: fused reverse step, prefetched 3-channel scatter, device-scope fix.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2c
timeout 120 python -m pytest tests/test_tma_gather.py -m gpu -q -p no:cacheprovider --timeout 100 > ${o}_tma.log 2>&1
echo "tma rc=$?" >> ${o}_tma.log
timeout 200 python tools/precision_probe.py 128 > ${o}_precision128.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 300 python tools/tune_eval.py 256 base gather_tma=0 gather_seg=64 brick=0 base > ${o}_tune.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-paint-bench > ${o}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_tma -s 8 -c 2 -o ${o}_full_gather \
  python tools/one_eval.py 256 1 > ${o}_full_gather.log 2>&1
tail -n 3 ${o}_tma.log; tail -n 6 ${o}_pytest_gpu.log; cat ${o}_precision128.log | grep -v Warn; cat ${o}_tune.log; head -c 300 ${o}_bench.json
