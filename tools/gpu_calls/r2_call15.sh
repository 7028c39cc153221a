#!/bin/bash
# Round 2, GPU call 15: whole GPU suite and bench line with the fused bias passes, the fused redshift-space paint and the
# streaming 3-channel scatter.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2o
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
tail -n 8 ${o}_pytest_gpu.log; head -c 600 ${o}_bench.json; tail -n 3 ${o}_bench_err.log
