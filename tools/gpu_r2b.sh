#!/bin/bash
# Round-2 diagnostics: latency of the network kernel at small batches (per-layer trace + per-launch time), new tests.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b
mkdir -p "$OUT"
for n in 2 100 296 592 1184; do
  timeout 120 python tools/net_bench.py --n $n --reps 200 >> "$OUT/net_small.jsonl" 2>> "$OUT/net_small.err"
done
timeout 120 python tools/net_trace.py 10 128 2 > "$OUT/trace_n2.txt" 2>&1
timeout 120 python tools/net_trace.py 10 128 100 > "$OUT/trace_n100.txt" 2>&1
timeout 120 python tools/net_trace.py 10 128 592 > "$OUT/trace_n592.txt" 2>&1
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short -s > "$OUT/pytest_gpu.log" 2>&1; echo "pytest all rc=$?" | tee "$OUT/summary.txt"
cat "$OUT/net_small.jsonl"; head -30 "$OUT/trace_n2.txt"; tail -4 "$OUT/trace_n100.txt"; tail -4 "$OUT/trace_n592.txt"
grep -n "single-board\|passed\|failed" "$OUT/pytest_gpu.log" | tail -5; tail -30 "$OUT/pytest_gpu.log" | grep -v "^$" | tail -25
