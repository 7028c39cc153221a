#!/bin/bash
# Round 2, GPU call 17 (TWO B200s): the bench's slab arm with the halo sized from the warm-up evaluation vs the fixed 24 planes.
set -u
mkdir -p gpurun_out
o=gpurun_out/r2q
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 400 bash -c "$(declare -f run); run 29631 bench.py --gpus 2 --steps 5 --warmup 3" > ${o}_bench2.json 2> ${o}_bench2_err.log
echo "bench2 rc=$?" >> ${o}_bench2_err.log
timeout 400 bash -c "$(declare -f run); run 29632 bench.py --gpus 2 --steps 5 --warmup 3 --fixed-halo" > ${o}_bench2_fixed.json 2> ${o}_bench2_fixed_err.log
echo "bench2 fixed rc=$?" >> ${o}_bench2_fixed_err.log
for f in bench2 bench2_fixed; do python - <<PY
import json
for line in open("${o}_${f}.json"):
    if line.startswith("{"):
        d = json.loads(line); print("${f}", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("slab", {}).get("halo_planes"), d.get("slab", {}).get("halo_note"))
PY
tail -n 3 ${o}_${f}_err.log; done
