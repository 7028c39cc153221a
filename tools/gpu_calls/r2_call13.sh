#!/bin/bash
# Round 2, GPU call 13: streaming brick scatter after the instruction diet -- parity, timing, ncu.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2m
timeout 300 python -m pytest tests/test_brick_stream.py "tests/test_abi_parity.py::test_brick_scatter_matches_generic" -m gpu -q -p no:cacheprovider --timeout 200 > ${o}_pytest.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest.log
timeout 200 python tools/brick_probe.py 256 > ${o}_probe.log 2>&1
timeout 300 python tools/tune_eval.py 256 base brick_stream=0 base > ${o}_tune.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:brick_stream_kernel -s 12 -c 2 -o ${o}_full_stream \
  python tools/brick_probe.py 256 > ${o}_ncu.log 2>&1
tail -n 8 ${o}_pytest.log; cat ${o}_probe.log | tail -n 12; cat ${o}_tune.log
