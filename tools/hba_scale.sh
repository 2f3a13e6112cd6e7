#!/bin/bash
# hierarchical-GBA leg of bench.py at N GPUs (N = $1): prints the gba block compactly
N=$1
if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py"; fi
timeout 600 $CMD --gpus $N --steps 3 --no-cpu-baseline --no-local-mapping $HBA_ARGS > gpurun_out/hba_n$N.json 2> gpurun_out/hba_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/hba_n$N.json").read().strip().splitlines()[-1])
g=d["gba"]
print("N=$N value", d["value"], "| hba ms/pass", round(g.get("ms_per_step",-1),1), "passes/s", round(g.get("value",-1),2), g.get("error",""))
for r in g.get("per_rank_[bottom,merge,exchange,top,wall]_ms", []): print("   ", r)
print("  bottom:", g.get("bottom"))
print("  rank0 per step:", g.get("per_step_[bottom,merge,exchange,top,wall]_ms_rank0"))
PY
