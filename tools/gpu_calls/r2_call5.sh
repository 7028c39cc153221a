#!/bin/bash
# Round 2, GPU call 5: fused (y,z) FFT kernels, brick-ordered gather mapping.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2e
timeout 200 python -m pytest tests/test_yzfft.py tests/test_tma_gather.py -m gpu -q -p no:cacheprovider --timeout 150 > ${o}_new.log 2>&1
echo "new rc=$?" >> ${o}_new.log
timeout 300 python tools/tune_eval.py 256 base yzfft=0 gather_brick=0 base > ${o}_tune.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-paint-bench > ${o}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:c_yz_kernel -s 6 -c 2 -o ${o}_full_yz \
  python tools/one_eval.py 256 1 > ${o}_full_yz.log 2>&1
tail -n 15 ${o}_new.log; cat ${o}_tune.log; tail -n 6 ${o}_pytest_gpu.log; head -c 300 ${o}_bench.json
