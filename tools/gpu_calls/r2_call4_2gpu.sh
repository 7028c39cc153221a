#!/bin/bash
# Round 2, GPU call 4 (TWO B200s): the slab-decomposed bench arm under torchrun, as the driver launches it.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/r2_call4_2gpu.sh'
set -u
mkdir -p gpurun_out
o=gpurun_out/r2d
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 5 --warmup 3 > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  tools/slab_bench.py --mesh 256 --steps 3 --warmup 2 --model --model-check > ${o}_slab2_check.json 2> ${o}_slab2_check_err.log
echo "slab2 rc=$?" >> ${o}_slab2_check_err.log
head -c 1500 ${o}_bench2.json; echo; tail -n 5 ${o}_bench2_err.log; head -c 800 ${o}_slab2_check.json; tail -n 3 ${o}_slab2_check_err.log
