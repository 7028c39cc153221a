#!/bin/bash
# Round 2, GPU call 29 (EIGHT B200s): BASELINE C5 (1024^3 mesh / particles, 2048^3 paint mesh, 20 steps + adjoint, force
# meshes recomputed) with the per-step active halo planes; the bench's slab arm once more on the final code.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2ac
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 420 bash -c "$(declare -f run); run 29711 tools/slab_bench.py --mesh 1024 --steps 2 --warmup 1 --nbody-steps 20 --model --oversamp 2 --no-force-tape --auto-halo" > ${o}_c5.json 2> ${o}_c5_err.log
echo "c5 rc=$?" >> ${o}_c5_err.log
timeout 300 bash -c "$(declare -f run); run 29712 bench.py --gpus 8 --steps 10 --warmup 3" > ${o}_bench8.json 2> ${o}_bench8_err.log
echo "bench8 rc=$?" >> ${o}_bench8_err.log
grep "^{" ${o}_c5.json | head -c 1200; echo; grep -v "^\s*\^*$" ${o}_c5_err.log | tail -n 4; head -c 300 ${o}_bench8.json; echo; tail -n 2 ${o}_bench8_err.log
