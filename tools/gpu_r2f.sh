#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f
mkdir -p "$OUT"
R="$OUT/variants.txt"
echo "== pair engine, parallel relay" >> $R
for n in 2 100 592 18944; do timeout 120 python tools/net_bench.py --n $n --reps 40 --engine tcgen05_pair >> $R 2>&1; done
timeout 120 python tools/net_trace.py 10 128 592 tcgen05_pair 2>&1 | sed -n 4,7p >> $R
echo "== pair one-tile for tiny" >> $R
OTH_LATENCY_SHAPE_PAIR=1 timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 4,7p >> $R
for n in 2 100 296; do OTH_LATENCY_SHAPE_PAIR=1 timeout 120 python tools/net_bench.py --n $n --reps 100 >> $R 2>&1; done
echo "== tests pair" >> $R
OTH_LATENCY_SHAPE_PAIR=1 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -3 >> $R
echo "== tests CL=2" >> $R
OTH_TC_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -30 >> $R
cat $R
