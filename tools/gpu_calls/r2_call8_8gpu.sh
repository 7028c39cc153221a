#!/bin/bash
# Round 2, GPU call 8 (EIGHT B200s of one box): the bench arm as the driver's scaling run launches it (weak scaling:
# 512^3 on 8 GPUs, 256^3 cells per rank), the 512^3 strong-scaling point of round 1's table, and BASELINE C5
# (1024^3 mesh / particles, 2x paint oversampling = 2048^3 paint mesh, 20 steps + adjoint).
set -u
mkdir -p gpurun_out
o=gpurun_out/r2h
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 bash -c "$(declare -f run); run 29621 bench.py --gpus 8 --steps 5 --warmup 3" > ${o}_bench8.json 2> ${o}_bench8_err.log
echo "bench8 rc=$?" >> ${o}_bench8_err.log
timeout 300 bash -c "$(declare -f run); run 29622 tools/slab_bench.py --mesh 512 --steps 3 --warmup 2 --model" > ${o}_slab8_512.json 2> ${o}_slab8_512_err.log
echo "slab512 rc=$?" >> ${o}_slab8_512_err.log
timeout 500 bash -c "$(declare -f run); run 29623 tools/slab_bench.py --mesh 1024 --steps 2 --warmup 1 --nbody-steps 20 --model --oversamp 2 --no-force-tape" > ${o}_c5_recompute.json 2> ${o}_c5_recompute_err.log
echo "c5 recompute rc=$?" >> ${o}_c5_recompute_err.log
timeout 500 bash -c "$(declare -f run); run 29624 tools/slab_bench.py --mesh 1024 --steps 2 --warmup 1 --nbody-steps 20 --model --oversamp 2" > ${o}_c5_taped.json 2> ${o}_c5_taped_err.log
echo "c5 taped rc=$?" >> ${o}_c5_taped_err.log
for f in bench8 slab8_512 c5_recompute c5_taped; do grep "^{" ${o}_${f}.json | head -c 900; echo; tail -n 2 ${o}_${f}_err.log; done
