#!/bin/bash
# hierarchical-GBA leg ALONE (bench.py --workload hba) at N GPUs ($1), extra env in $2..: prints phases
N=$1; shift
if [ "$N" = "1" ]; then CMD="python bench.py"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py"; fi
env "$@" timeout 600 $CMD --gpus $N --workload hba $HBA_ARGS > gpurun_out/hbao_n$N.json 2> gpurun_out/hbao_n$N.err
python - <<PY
import json
g=json.loads(open("gpurun_out/hbao_n$N.json").read().strip().splitlines()[-1])
print("N=$N $*", "| ms/pass", round(g.get("ms_per_step",-1),1), "passes/s", round(g.get("value",-1),2), g.get("error",""))
for r in g.get("per_rank_[bottom,merge,exchange,top,wall]_ms", [])[:2]: print("   ", r)
print("  rank0 per step:", g.get("per_step_[bottom,merge,exchange,top,wall]_ms_rank0"))
PY
