#!/bin/bash
# Tensor-map TMA (UTMALDG) vs 1-D bulk copies (UBLKCP) for the trunk's weight stream.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2h
mkdir -p "$OUT"
R="$OUT/variants.txt"
run() {
  echo "== $*" >> $R
  ( export "$@" DUMMY=1
    timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 4,6p >> $R
    timeout 120 python tools/net_trace.py 10 128 592 2>&1 | sed -n 4,6p >> $R
    for n in 100 592 18944; do timeout 120 python tools/net_bench.py --n $n --reps 40 >> $R 2>&1; done )
}
run A=base
run OTH_TC_TMAP=1
run OTH_TC_TMAP=1 OTH_TC_ONE_TILE=1
run OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so OTH_TC_TMAP=1
run OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so OTH_TC_TMAP=1 OTH_TC_ONE_TILE=1
echo "== tests TMAP" >> $R
OTH_TC_TMAP=1 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -12 >> $R
echo "== tests TMAP g2 one-tile" >> $R
OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so OTH_TC_TMAP=1 OTH_TC_ONE_TILE=1 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -12 >> $R
cat $R
