#!/bin/bash
# First-light run on a B200 box: every stage in its own process with its own timeout, logs under
# gpurun_out/, never stops at the first failure (we want all the evidence from one box lease).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/firstlight
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,driver_version --format=csv > "$OUT/gpu.txt" 2>&1
python - <<'PY' > "$OUT/import.log" 2>&1
import othello_reinforcement_learning_test_b200 as p
print(p._lib.load().oth_version())
ctx = p.Context.default(0)
print("ctx ok")
PY
echo "import rc=$?" | tee -a "$OUT/summary.txt"
timeout 300 python tests/oth_umma_probe.py > "$OUT/umma_probe.log" 2>&1
echo "umma_probe rc=$?" | tee -a "$OUT/summary.txt"
for t in a_bitboard b_search c_selfplay d_net_simt e_umma_probe f_net_tc g_end_to_end; do
  timeout 900 python -m pytest tests/test_gpu_${t}.py -q -m gpu -x --tb=short > "$OUT/test_${t}.log" 2>&1
  echo "test_gpu_${t} rc=$?" | tee -a "$OUT/summary.txt"
done
tail -n 5 "$OUT"/test_*.log
cat "$OUT/umma_probe.log"
cat "$OUT/summary.txt"
