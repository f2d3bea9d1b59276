#!/bin/bash
# Standard GPU round: full gpu test-suite, smoke, bench, ncu launch list + one full capture of the top kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/round
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -q -m gpu -x --tb=short > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?" | tee -a "$OUT/summary.txt"
if [ "${NCU:-1}" = "1" ]; then
  # launch list of the bench's own workload (one step, no warm-up: not a bench value): kernel SHARES are comparable
  # with roofline.kernel_share_of_step
  # (ncu costs ~45 ms per intercepted launch even when it only skips it, so the window sits early in a 37,888-game
  #  campaign and the application is killed once the window is captured)
  LL_CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --games 37888"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kill on -s 1500 -c 600 --csv --log-file "$OUT/launches.csv" $LL_CMD > "$OUT/ncu_launches.log" 2>&1
  echo "ncu launches rc=$?" | tee -a "$OUT/summary.txt"
  NCU_CMD="python bench.py --steps 1 --warmup 0 --games 2368 --no-cpu-baseline"
  timeout 600 $NCU_CMD > "$OUT/ncu_plain.log" 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 400 -c 2 -o "$OUT/prof_net_tc" -f $NCU_CMD > "$OUT/ncu_full.log" 2>&1
  echo "ncu full rc=$?" | tee -a "$OUT/summary.txt"
  timeout 300 python tools/net_bench.py --n 18944 --reps 20 > "$OUT/net_bench.json" 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_net_tc -s 3 -c 1 -o "$OUT/prof_net_tc_fullbatch" -f python tools/net_bench.py --n 18944 --reps 2 > "$OUT/ncu_net_bench.log" 2>&1
  echo "ncu net_bench rc=$?" | tee -a "$OUT/summary.txt"
  timeout 120 python tools/playout_bench.py 4194304 > "$OUT/playout_bench.json" 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_playouts -s 1 -c 1 -o "$OUT/prof_playouts" -f python tools/playout_bench.py 4194304 > "$OUT/ncu_playouts.log" 2>&1
  echo "ncu playouts rc=$?" | tee -a "$OUT/summary.txt"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_tree_select|k_tree_expand|k_tree_assign|k_playouts" -s 300 -c 8 -o "$OUT/prof_tree" -f $NCU_CMD > "$OUT/ncu_tree.log" 2>&1
  echo "ncu tree rc=$?" | tee -a "$OUT/summary.txt"
fi
tail -n 6 "$OUT/pytest_gpu.log"; cat "$OUT/smoke.log" | tail -3; cat "$OUT/bench.json"; tail -5 "$OUT/bench.err"; cat "$OUT/summary.txt"
