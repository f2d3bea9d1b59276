#!/bin/bash
# Is the trunk's weight stream bound by ring depth (latency) or by the per-SM copy rate?  Same kernel, ring of 6 / 3 / 2 slots,
# and 4 bulk copies per stage; one CTA (n=2) and a full batch.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c
mkdir -p "$OUT"
for tag in base s3 s2 c4; do
  if [ "$tag" = base ]; then unset OTH_LIB_PATH; else export OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_$tag.so; fi
  echo "== $tag" >> "$OUT/variants.txt"
  timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 4,8p >> "$OUT/variants.txt"
  timeout 120 python tools/net_bench.py --n 2 --reps 200 >> "$OUT/variants.txt" 2>&1
  timeout 120 python tools/net_bench.py --n 18944 --reps 20 >> "$OUT/variants.txt" 2>&1
done
cat "$OUT/variants.txt"
