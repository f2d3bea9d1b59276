#!/bin/bash
# Round-2 first GPU pass: new parity tests, then quick small-campaign numbers for both schedules, then the full suite.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,driver_version --format=csv > "$OUT/gpu.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_c_selfplay.py tests/test_gpu_g_end_to_end.py tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short > "$OUT/pytest_new.log" 2>&1; echo "pytest new rc=$?" | tee "$OUT/summary.txt"
timeout 600 python bench.py --steps 1 --warmup 1 --games 18944 --warmup-games 2368 --cpu-seconds 5 > "$OUT/bench_small.json" 2> "$OUT/bench_small.err"; echo "bench small rc=$?" | tee -a "$OUT/summary.txt"
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > "$OUT/pytest_gpu.log" 2>&1; echo "pytest all rc=$?" | tee -a "$OUT/summary.txt"
tail -n 15 "$OUT/pytest_new.log"; tail -n 5 "$OUT/pytest_gpu.log"; tail -5 "$OUT/bench_small.err"; cat "$OUT/summary.txt"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2a/bench_small.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "wall", d.get("wall_s_total"))
    for k, v in d.get("legs", {}).items():
        print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("what",)})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("no bench json", e)
PY
