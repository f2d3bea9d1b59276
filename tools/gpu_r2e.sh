#!/bin/bash
# Weight-stream experiments on k_net_tc: 16 KB stage groups (g2), cluster multicast (CL = 2, 4), one tile per CTA.
# (Kept as the record of how profiles/r02_weight_stream_experiments_1.txt was produced; the OTH_TC_CLUSTER* knobs belonged to the
#  cluster-multicast experiment, which was removed from net_tc.cu afterwards -- see DESIGN.md section 4.)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e
mkdir -p "$OUT"
R="$OUT/variants.txt"
run() {  # tag, env...
  echo "== $*" >> $R
  ( export "$@" DUMMY=1
    timeout 120 python tools/net_trace.py 10 128 2 2>&1 | sed -n 4,6p >> $R
    for n in 100 592 18944; do timeout 120 python tools/net_bench.py --n $n --reps 40 >> $R 2>&1; done )
}
run A=base
run OTH_TC_CLUSTER=2
run OTH_TC_CLUSTER=4
run OTH_TC_ONE_TILE=1 OTH_TC_CLUSTER_SMALL=2
run OTH_TC_ONE_TILE=1 OTH_TC_CLUSTER_SMALL=4
run OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so
run OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so OTH_TC_CLUSTER=2
echo "== tests CL=2" >> $R
OTH_TC_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -3 >> $R
echo "== tests CL=4 one tile small" >> $R
OTH_TC_ONE_TILE=1 OTH_TC_CLUSTER_SMALL=4 OTH_TC_CLUSTER=2 timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -3 >> $R
echo "== tests g2" >> $R
OTH_LIB_PATH=$PWD/build_tmp/libothello_b200_g2.so timeout 300 python -m pytest tests/test_gpu_f_net_tc.py -q -m gpu -x --tb=short 2>&1 | tail -3 >> $R
cat $R
