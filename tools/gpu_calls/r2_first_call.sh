#!/bin/bash
# First GPU call of round 2 (one B200): settles what round 1 could not re-measure, then refreshes the headline numbers.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
# Every step has its own timeout and writes into gpurun_out/; nothing here runs under a profiler except the last step.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2a
# 1. the discriminating cross-checks at 256^3 (Hermitian projection before C2R) -- see profiles/r1_slab_model_1gpu.md
timeout 200 python -m pytest tests/test_zz_cross_check_256.py -m gpu -q -p no:cacheprovider > ${o}_cross_check.log 2>&1
echo "cross-check rc=$?" >> ${o}_cross_check.log
# 2. the same comparison with its numbers, at the benchmark cell size and at 10 Mpc/h cells, and with a 2x paint mesh
for args in "--cell 2.5" "--cell 10" "--cell 2.5 --oversamp 2"; do
  tag=$(echo $args | tr -d ' -.')
  timeout 120 python tools/slab_bench.py --mesh 256 --steps 3 --warmup 1 --model --model-check $args \
    > ${o}_slab_model_${tag}.json 2> ${o}_slab_model_${tag}_err.log
done
# 3. the whole GPU suite, then the bench line (plain), then its launch list
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider > ${o}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest_gpu.log
timeout 300 python bench.py > ${o}_bench.json 2> ${o}_bench_err.log && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file ${o}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > ${o}_ncu.log 2>&1
# 4. A/B of the kernel variants built blind at the end of round 1 (whole evaluation, same inputs)
timeout 120 python tools/tune_eval.py 256 base brick_zmerge=1 base brick_zmerge=1 > ${o}_tune_zmerge.log 2>&1
# 5. per-operator timings, including the kernels that have parity tests but no timing yet
timeout 200 python tools/microbench.py --n 256 --out ${o}_microbench.json > ${o}_microbench.log 2>&1
tail -3 ${o}_cross_check.log ${o}_pytest_gpu.log; head -c 600 ${o}_bench.json
