#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2k
mkdir -p "$OUT"
R="$OUT/lat.txt"
timeout 900 python -m pytest tests/test_gpu_c_selfplay.py tests/test_gpu_g_end_to_end.py tests/test_gpu_i_fullsize.py -q -m gpu -x --tb=short > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$OUT/summary.txt"
tail -4 "$OUT/pytest.log" >> $R
for c in 3 4 6 8 12; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 100 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
for c in 2 4 6; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 4096 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
for c in 4 8; do OTH_ASYNC_MAX_STEPS=$c timeout 300 python tools/sched_bench.py --games 592 --schedule async --tag cap$c >> $R 2>> "$OUT/err.txt"; done
timeout 300 python tools/sched_bench.py --games 18944 --schedule async --reps 1 --tag auto >> $R 2>> "$OUT/err.txt"
timeout 300 python tools/sched_bench.py --games 18944 --schedule lockstep --reps 1 --tag auto >> $R 2>> "$OUT/err.txt"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file "$OUT/launches_g100.csv" python tools/sched_bench.py --games 100 --schedule async --reps 1 > "$OUT/ncu_g100.log" 2>&1
python - <<'PY' >> $R
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2k/launches_g100.csv")) if len(r) > 5]
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
