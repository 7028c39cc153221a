#!/bin/bash
# Round 2, GPU call 10 (EIGHT B200s): the 512^3 strong-scaling point of round 1's table with the whole model, and
# BASELINE C5 (1024^3 mesh / particles, 2x paint oversampling = 2048^3 paint mesh, 20 steps + adjoint) with the force
# meshes recomputed in the reverse sweep.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2j
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 240 bash -c "$(declare -f run); run 29622 tools/slab_bench.py --mesh 512 --steps 3 --warmup 2 --model" > ${o}_slab8_512.json 2> ${o}_slab8_512_err.log
echo "slab512 rc=$?" >> ${o}_slab8_512_err.log
timeout 420 bash -c "$(declare -f run); run 29623 tools/slab_bench.py --mesh 1024 --steps 2 --warmup 1 --nbody-steps 20 --model --oversamp 2 --no-force-tape" > ${o}_c5_recompute.json 2> ${o}_c5_recompute_err.log
echo "c5 recompute rc=$?" >> ${o}_c5_recompute_err.log
for f in slab8_512 c5_recompute; do grep "^{" ${o}_${f}.json | head -c 1500; echo; grep -v "^\s*\^*$" ${o}_${f}_err.log | tail -n 12; done
