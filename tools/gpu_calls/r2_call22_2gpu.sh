#!/bin/bash
# Round 2, GPU call 22 (TWO B200s): slab tests on the GPU and the 2-GPU bench line after the guard kernel.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2v
timeout 400 python -m pytest tests/test_dist_slab.py tests/test_zz_cross_check_256.py -m gpu -q -p no:cacheprovider --timeout 300 > ${o}_pytest.log 2>&1
echo "pytest rc=$?" >> ${o}_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 2 --steps 5 --warmup 3 > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29662 tools/slab_bench.py --mesh 256 --model --model-check --steps 3 --warmup 2 --check 128 > ${o}_slab2_check.json 2> ${o}_slab2_check_err.log
echo "slab2 rc=$?" >> ${o}_slab2_check_err.log
tail -n 4 ${o}_pytest.log; head -c 330 ${o}_bench2.json; echo; tail -n 2 ${o}_bench2_err.log; grep "^{" ${o}_slab2_check.json | head -c 900; echo; tail -n 2 ${o}_slab2_check_err.log
