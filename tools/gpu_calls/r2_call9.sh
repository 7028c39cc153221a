#!/bin/bash
# Round 2, GPU call 9: state of the suite and the bench line at head.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2i
timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -x > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 400 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
tail -n 6 ${o}_pytest_gpu.log; head -c 400 ${o}_bench.json; tail -n 3 ${o}_bench_err.log
