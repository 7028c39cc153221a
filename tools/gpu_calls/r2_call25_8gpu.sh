#!/bin/bash
# Round 2, GPU call 25 (EIGHT B200s): the bench's slab arm as the driver's scaling run launches it (512^3 on 8 GPUs).
set -u
mkdir -p gpurun_out
o=gpurun_out/r2y
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 8 --steps 5 --warmup 3 > ${o}_bench8.json 2> ${o}_bench8_err.log
echo "bench8 rc=$?" >> ${o}_bench8_err.log
head -c 400 ${o}_bench8.json; echo; tail -n 3 ${o}_bench8_err.log
