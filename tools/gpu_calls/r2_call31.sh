#!/bin/bash
# Round 2, GPU call 31 (the last 9.8 GPU-minutes): the whole GPU suite with the observation chain inside the paint, the
# two-field slab transforms and the emulated two-rank peer test; smoke; the fused-vs-elementwise probe; the bench line.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ae
timeout 330 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
tail -n 6 ${o}_pytest_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1
echo "smoke rc=$?" >> ${o}_smoke.log
tail -n 2 ${o}_smoke.log
timeout 120 python tools/obs_probe.py 128 > ${o}_obs_probe.log 2>&1
echo "probe rc=$?" >> ${o}_obs_probe.log
tail -n 3 ${o}_obs_probe.log
for two in 1 0; do
  MCPM_SLAB_TWO_FIELD=$two timeout 100 python tools/slab_bench.py --mesh 256 --model --model-check --steps 3 --warmup 2 \
    > ${o}_slab1_two${two}.json 2> ${o}_slab1_two${two}_err.log
  echo "two_field=$two rc=$?"; grep "^{" ${o}_slab1_two${two}.json | head -c 700; echo
done
timeout 240 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log
echo "bench rc=$?" >> ${o}_bench_err.log
head -c 400 ${o}_bench.json; echo; tail -n 2 ${o}_bench_err.log
