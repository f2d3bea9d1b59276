#!/bin/bash
# Latency shape: crossover batch size, small-campaign numbers, kernel durations inside a 100-game campaign (ncu launch list).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2j
mkdir -p "$OUT"
R="$OUT/lat.txt"
for n in 128 160 192 224 256; do
  OTH_LATENCY_SHAPE_MAX=296 timeout 120 python tools/net_bench.py --n $n --reps 200 | sed 's/^/lat  /' >> $R 2>&1
  OTH_NO_LATENCY_SHAPE=1 timeout 120 python tools/net_bench.py --n $n --reps 200 | sed 's/^/thr  /' >> $R 2>&1
done
timeout 600 python -m pytest tests/test_gpu_f_net_tc.py tests/test_gpu_g_end_to_end.py tests/test_gpu_c_selfplay.py -q -m gpu -x --tb=short > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -4 "$OUT/pytest.log" >> $R
for g in 100 256 592 1024 4096; do
  timeout 300 python tools/sched_bench.py --games $g --schedule async --tag lat >> $R 2>> "$OUT/err.txt"
  timeout 300 python tools/sched_bench.py --games $g --schedule lockstep --tag lat >> $R 2>> "$OUT/err.txt"
done
for c in 2 3 6; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 100 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file "$OUT/launches_g100.csv" python tools/sched_bench.py --games 100 --schedule async --reps 1 > "$OUT/ncu_g100.log" 2>&1
python - <<'PY' >> $R
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2j/launches_g100.csv")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
k = rows[hdr].index("Kernel Name"); v = rows[hdr].index("Metric Value"); u = rows[hdr].index("Metric Unit")
acc = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try:
        t = float(r[v].replace(",", "")); t = t / 1000.0 if r[u] == "ns" else t
        acc[r[k].split("(")[0][:40]].append(t)
    except Exception:
        pass
for name, ts in acc.items():
    print(f"ncu g100  {name:42s} n={len(ts):4d} mean={sum(ts)/len(ts):8.2f} us  max={max(ts):8.2f}")
PY
cat $R | cut -c 1-330; tail -3 "$OUT/err.txt"
