#!/bin/bash
# Round 2, GPU call 34 (the last two GPU-minutes): the GPU suite as the driver runs it, on the final tree.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ah
timeout 115 python -m pytest tests -x -q -m gpu -p no:cacheprovider > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
tail -n 5 ${o}_pytest_gpu.log
