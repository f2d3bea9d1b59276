#!/bin/bash
# The driver's own command lines: full GPU suite, then bench.py with --steps 20 --warmup 5 (both arms), wall clock recorded.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2g
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
S=$(date +%s)
timeout 1100 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$? wall=$(( $(date +%s) - S ))s" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_gpu.log"; tail -2 "$OUT/smoke.log"; tail -5 "$OUT/bench.err"; cat "$OUT/summary.txt"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2g/bench.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "schedule", "wall_s_total", "wall_s_timed_region")})
print("e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "clocks", d["clocks"])
print("tree", d["secondary_rooflines"]["tree_kernels"])
for k, v in d.get("legs", {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("what", "reference_signature", "packed", "note")})
print("cpu", d.get("cpu_baseline"))
PY
