#!/bin/bash
# Second GPU call of round 2 (two B200s): the slab-decomposed whole model under NCCL, which round 1 could only run on
# gloo / the CPU port and on one GPU.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_second_call_2gpu.sh'
set -u
mkdir -p gpurun_out
o=gpurun_out/r2b
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 "$@"; }
# parity of every rank's slab against the single-GPU FieldModel (computed on each GPU), then timing
timeout 150 bash -c "$(declare -f run); run tools/slab_bench.py --mesh 256 --steps 3 --warmup 1 --model --model-check --cell 10" \
  > ${o}_model256_cell10.json 2> ${o}_model256_cell10_err.log
timeout 150 bash -c "$(declare -f run); run tools/slab_bench.py --mesh 256 --steps 3 --warmup 1 --model --model-check --oversamp 2" \
  > ${o}_model256_over2.json 2> ${o}_model256_over2_err.log
timeout 200 bash -c "$(declare -f run); run tools/slab_bench.py --mesh 512 --steps 3 --warmup 1 --model --model-check --model-check-mesh 128" \
  > ${o}_model512.json 2> ${o}_model512_err.log
tail -c 700 ${o}_model256_cell10.json ${o}_model256_over2.json ${o}_model512.json
