#!/bin/bash
# Round 2, GPU call 28: nufft VJP with the interlaced transposes in one gather -- parity tests, evaluation time.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ab
timeout 500 python -m pytest tests/test_abi_parity.py tests/test_api_model.py tests/test_full_size_oracle.py tests/test_metrics_parity.py -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest.log
timeout 300 python tools/tune_eval.py 256 base base > ${o}_tune.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-paint-bench > ${o}_bench.json 2> ${o}_bench_err.log
tail -n 4 ${o}_pytest.log; cat ${o}_tune.log; head -c 250 ${o}_bench.json
