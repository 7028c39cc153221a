#!/bin/bash
# Round 2, GPU call 11: the streaming brick scatter -- parity, per-kernel timing, whole-evaluation A/B.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2k
timeout 300 python -m pytest tests/test_brick_stream.py "tests/test_abi_parity.py::test_brick_scatter_matches_generic" tests/test_full_size.py -m gpu -q -p no:cacheprovider --timeout 200 -x > ${o}_pytest.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest.log
timeout 200 python tools/brick_probe.py 256 > ${o}_probe.log 2>&1
timeout 300 python tools/tune_eval.py 256 base brick_stream=0 brick_stream=48 base > ${o}_tune.log 2>&1
tail -n 12 ${o}_pytest.log; cat ${o}_probe.log | tail -n 12; cat ${o}_tune.log
