#!/bin/bash
# Round 2, GPU call 33 (the last GPU-minutes): the GPU suite and smoke on the final tree (light-cone pass, displacements
# carried into the observed paint), and the fused-vs-elementwise probe at 128^3 and 256^3.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ag
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 100 > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
tail -n 5 ${o}_pytest_gpu.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1
echo "smoke rc=$?" >> ${o}_smoke.log
tail -n 2 ${o}_smoke.log
timeout 40 python tools/obs_probe.py 128 > ${o}_obs_probe_128.log 2>&1
echo "probe rc=$?" >> ${o}_obs_probe_128.log
tail -n 3 ${o}_obs_probe_128.log
timeout 80 python tools/obs_probe.py 256 > ${o}_obs_probe_256.log 2>&1
echo "probe rc=$?" >> ${o}_obs_probe_256.log
tail -n 3 ${o}_obs_probe_256.log
