#!/bin/bash
# Round 2, GPU call 32 (TWO B200s, the last GPU-minutes): the slab model with the two-field distributed x-transform
# against the single-GPU model and against the three-field exchange; then the driver's 2-GPU bench command.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2af
port=29701
for two in 1 0; do
  MCPM_SLAB_TWO_FIELD=$two timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port $port tools/slab_bench.py --mesh 256 --model --model-check --steps 5 --warmup 2 --auto-halo \
    > ${o}_slab2_two${two}.json 2> ${o}_slab2_two${two}_err.log
  echo "two_field=$two rc=$?"; grep "^{" ${o}_slab2_two${two}.json | head -c 900; echo
  port=$((port + 1))
done
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --steps 10 --warmup 3 > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
head -c 330 ${o}_bench2.json; echo; tail -n 2 ${o}_bench2_err.log
