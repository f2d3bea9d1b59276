#!/bin/bash
# Latency shape k_net_lat: parity (bit identity with the throughput kernel), per-layer trace, launch latency, small campaigns.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2i
mkdir -p "$OUT"
R="$OUT/lat.txt"
timeout 600 python -m pytest tests/test_gpu_f_net_tc.py tests/test_gpu_d_net_simt.py tests/test_gpu_g_end_to_end.py -q -m gpu -x --tb=short > "$OUT/pytest_net.log" 2>&1; echo "pytest net rc=$?" | tee "$OUT/summary.txt"
tail -15 "$OUT/pytest_net.log" >> $R
timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 1,12p >> $R
timeout 120 python tools/net_trace.py 5 64 2 2>&1 | sed -n 4,8p >> $R
for n in 2 100 296 298; do timeout 120 python tools/net_bench.py --n $n --reps 200 >> $R 2>&1; done
OTH_NO_LATENCY_SHAPE=1 timeout 120 python tools/net_bench.py --n 100 --reps 200 >> $R 2>&1
timeout 120 python tools/net_bench.py --n 100 --reps 200 --blocks 5 --filters 64 >> $R 2>&1
OTH_NO_LATENCY_SHAPE=1 timeout 120 python tools/net_bench.py --n 100 --reps 200 --blocks 5 --filters 64 >> $R 2>&1
timeout 300 python tools/sched_bench.py --games 100 --schedule async --tag lat >> $R 2>> "$OUT/err.txt"
timeout 300 python tools/sched_bench.py --games 100 --schedule lockstep --tag lat >> $R 2>> "$OUT/err.txt"
timeout 300 python tools/sched_bench.py --games 256 --schedule async --tag lat >> $R 2>> "$OUT/err.txt"
OTH_NO_LATENCY_SHAPE=1 timeout 300 python tools/sched_bench.py --games 256 --schedule async --tag r1net >> $R 2>> "$OUT/err.txt"
cat $R | cut -c 1-400; tail -5 "$OUT/err.txt"
