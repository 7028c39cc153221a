#!/bin/bash
# Round 2, GPU call 30: driver-style closing run -- build check, smoke, the whole GPU suite, the bench line.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ad
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1
echo "smoke rc=$?" >> ${o}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 500 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
tail -n 3 ${o}_smoke.log; tail -n 4 ${o}_pytest_gpu.log; head -c 300 ${o}_bench.json; echo; tail -n 2 ${o}_bench_err.log
