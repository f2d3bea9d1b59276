#!/bin/bash
# Multi-GPU arm exactly as the driver launches it (torchrun, one rank per GPU), short.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}; STEPS=${2:-2}; WARM=${3:-1}
OUT=gpurun_out/mg_r2_$N
mkdir -p "$OUT"
S=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $STEPS --warmup $WARM > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "rc=$? wall=$(( $(date +%s) - S ))s" | tee "$OUT/summary.txt"
tail -5 "$OUT/bench.err"
python - "$OUT/bench.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "gpu_launches", "wall_s_total")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "collectives", d.get("collectives"))
print("roof", d["roofline"]["achieved"], "clocks", d["clocks"])
PY
