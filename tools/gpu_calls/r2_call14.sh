#!/bin/bash
# Round 2, GPU call 14: streaming brick scatter after the per-brick broadcast, fused lagrangian_bias passes on the GPU.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2n
timeout 400 python -m pytest tests/test_brick_stream.py "tests/test_abi_parity.py::test_brick_scatter_matches_generic" tests/test_api_model.py -k "brick or bias or general_evolve or field_level" -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest.log
timeout 200 python tools/brick_probe.py 256 > ${o}_probe.log 2>&1
timeout 300 python tools/tune_eval.py 256 base brick_stream=0 base brick_stream=0 > ${o}_tune.log 2>&1
tail -n 8 ${o}_pytest.log; cat ${o}_probe.log | tail -n 12; cat ${o}_tune.log
