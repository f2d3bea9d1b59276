#!/bin/bash
# Round-2 ncu evidence.  Every command is first run plain (exit 0) and then under ncu; numbers printed under ncu are never bench values.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/prof_r2
mkdir -p "$OUT"
# 1) launch list of a bench campaign (37,888 games, window in the early middle game)
LL_CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-legs --games 37888"
timeout 400 $LL_CMD > "$OUT/ll_plain.json" 2> "$OUT/ll_plain.err"; echo "ll plain rc=$?" | tee "$OUT/summary.txt"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kill on -s 1500 -c 600 --csv --log-file "$OUT/launches.csv" $LL_CMD > "$OUT/ncu_launches.log" 2>&1
echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
# 2) the throughput kernel on a full resident batch, 3) the latency shape on 100 positions
timeout 300 python tools/net_bench.py --n 18944 --reps 20 > "$OUT/net_bench.json" 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 3 -c 1 -o "$OUT/prof_net_tc_fullbatch" -f python tools/net_bench.py --n 18944 --reps 2 > "$OUT/ncu_net_bench.log" 2>&1
echo "ncu net_tc rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python tools/net_bench.py --n 100 --reps 200 > "$OUT/net_lat_bench.json" 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_net_lat -s 3 -c 1 -o "$OUT/prof_net_lat" -f python tools/net_bench.py --n 100 --reps 2 > "$OUT/ncu_net_lat.log" 2>&1
echo "ncu net_lat rc=$?" | tee -a "$OUT/summary.txt"
# 4) tree kernels at bench-sized launches: a 151,552-game lock-step campaign, launches taken around ply 8
TREE_CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-legs --games 151552"
timeout 900 ncu --set full --clock-control none --import-source on --kill on -k "regex:k_tree_select|k_tree_expand|k_tree_assign" -s 1200 -c 6 -o "$OUT/prof_tree" -f $TREE_CMD > "$OUT/ncu_tree.log" 2>&1
echo "ncu tree rc=$?" | tee -a "$OUT/summary.txt"
# 5) the async advance kernel inside a 4,096-game campaign
timeout 600 ncu --set full --clock-control none --import-source on --kill on -k "regex:k_as_advance" -s 600 -c 2 -o "$OUT/prof_async" -f python tools/sched_bench.py --games 4096 --schedule async --reps 1 > "$OUT/ncu_async.log" 2>&1
echo "ncu async rc=$?" | tee -a "$OUT/summary.txt"
ls -la "$OUT"; cat "$OUT/summary.txt"
