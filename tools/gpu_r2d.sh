#!/bin/bash
# Latency shape of the network (CTA pairs, one tile per CTA) and the schedules at small campaigns.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_f_net_tc.py tests/test_gpu_d_net_simt.py -q -m gpu -x --tb=short > "$OUT/pytest_net.log" 2>&1; echo "pytest net rc=$?" | tee "$OUT/summary.txt"
for n in 2 100 296 400 592; do timeout 120 python tools/net_bench.py --n $n --reps 200 >> "$OUT/net_small.jsonl" 2>> "$OUT/err.txt"; done
OTH_NO_LATENCY_SHAPE=1 timeout 120 python tools/net_bench.py --n 100 --reps 200 >> "$OUT/net_small.jsonl" 2>> "$OUT/err.txt"
timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 1,10p > "$OUT/trace_n2.txt"
R="$OUT/sched.jsonl"
timeout 300 python tools/sched_bench.py --games 100 --schedule lockstep --tag lockstep >> $R 2>> "$OUT/err.txt"
OTH_NO_LATENCY_SHAPE=1 timeout 300 python tools/sched_bench.py --games 100 --schedule lockstep --tag lockstep_r1net >> $R 2>> "$OUT/err.txt"
for c in 2 4 6 12 104; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 100 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
timeout 300 python tools/sched_bench.py --games 4096 --schedule lockstep --tag lockstep >> $R 2>> "$OUT/err.txt"
for c in 2 4 6 12; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 4096 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
timeout 300 python tools/sched_bench.py --games 592 --schedule lockstep --tag lockstep >> $R 2>> "$OUT/err.txt"
OTH_ASYNC_MAX_STEPS=6 timeout 300 python tools/sched_bench.py --games 592 --schedule async --tag cap6 >> $R 2>> "$OUT/err.txt"
timeout 300 python tools/sched_bench.py --games 18944 --schedule lockstep --reps 1 --tag lockstep >> $R 2>> "$OUT/err.txt"
OTH_ASYNC_MAX_STEPS=6 timeout 300 python tools/sched_bench.py --games 18944 --schedule async --reps 1 --tag cap6 >> $R 2>> "$OUT/err.txt"
tail -3 "$OUT/pytest_net.log"; cat "$OUT/net_small.jsonl"; cat "$OUT/trace_n2.txt"; cat $R | cut -c 1-420; tail -5 "$OUT/err.txt"
