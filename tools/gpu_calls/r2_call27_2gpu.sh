#!/bin/bash
# Round 2, GPU call 27 (TWO B200s): the bench's slab arm after sizing the halo from all of its white fields.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2aa
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29691 bench.py --gpus 2 --steps 10 --warmup 3 > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29692 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > ${o}_ref2.json 2> ${o}_ref2_err.log
echo "ref2 rc=$?" >> ${o}_ref2_err.log
head -c 330 ${o}_bench2.json; echo; tail -n 2 ${o}_bench2_err.log; head -c 300 ${o}_ref2.json; echo; tail -n 2 ${o}_ref2_err.log
