#!/bin/bash
# Round 2, GPU call 16: the round's closing measurements -- suite subset touched since call 15, the bench line with its
# CPU legs, the reference arm, the launch list of the bench command and ncu --set full of the two top kernels.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2p
timeout 400 python -m pytest tests/test_brick_stream.py tests/test_abi_parity.py tests/test_api_model.py tests/test_full_size.py -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 500 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > ${o}_reference.json 2> ${o}_reference_err.log
echo "reference rc=$?" >> ${o}_reference_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-paint-bench > ${o}_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"brick_stream_kernel|gather_tma_kernel" -s 30 -c 3 -o ${o}_full_top \
  python tools/one_eval.py 256 1 > ${o}_full_top.log 2>&1
tail -n 4 ${o}_pytest_gpu.log; head -c 300 ${o}_bench.json; echo; head -c 300 ${o}_reference.json; echo; tail -n 2 ${o}_bench_err.log ${o}_reference_err.log ${o}_full_top.log
