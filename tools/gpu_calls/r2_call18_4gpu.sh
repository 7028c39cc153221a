#!/bin/bash
# Round 2, GPU call 18 (FOUR B200s): the bench's slab arm as the driver's scaling run launches it (512x512x256 on 4 GPUs).
set -u
mkdir -p gpurun_out
o=gpurun_out/r2r
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 4 --steps 5 --warmup 3 > ${o}_bench4.json 2> ${o}_bench4_err.log
echo "bench4 rc=$?" >> ${o}_bench4_err.log
head -c 400 ${o}_bench4.json; echo; tail -n 3 ${o}_bench4_err.log
