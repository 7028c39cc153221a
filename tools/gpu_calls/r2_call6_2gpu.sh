#!/bin/bash
# Round 2, GPU call 6 (TWO B200s): peer-memory halo kernels; section profile of the slab model; the bench arm.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2g
MCPM_SLAB_PROFILE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  tools/slab_bench.py --mesh 256 --steps 3 --warmup 3 --model --model-check --check 128 > ${o}_slab2_check.json 2> ${o}_slab2_check_err.log
echo "slab2 rc=$?" >> ${o}_slab2_check_err.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus 2 --steps 5 --warmup 3 > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
MCPM_SLAB_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
  tools/slab_bench.py --mesh 256 --steps 3 --warmup 3 --model > ${o}_slab2_nccl.json 2> ${o}_slab2_nccl_err.log
grep "^{" ${o}_slab2_check.json | head -c 2500; echo; tail -n 3 ${o}_slab2_check_err.log; grep "^{" ${o}_bench2.json | head -c 600; echo; tail -n 3 ${o}_bench2_err.log; grep "^{" ${o}_slab2_nccl.json | head -c 400
