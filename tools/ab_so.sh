#!/bin/bash
# Same-box A/B of two builds of the library: bash tools/ab_so.sh <a.so> <b.so> [rounds]
# Alternates the in-tree libhtrvt_b200.so between the two files and prints step time + the GEMM entries of the breakdown.
A=$1; B=$2; R=${3:-2}
for i in $(seq $R); do
  for v in A B; do
    f=$A; [ $v = B ] && f=$B
    cp $f htr-vt_b200/libhtrvt_b200.so
    python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
b=d['breakdown_ms']
print('$v', '%.3f ms/step' % d['ms_per_step'], d['clocks']['sm_mhz'], ' '.join('%s=%.3f' % (k, b[k]) for k in ('conv_fwd','conv_dgrad_bn','conv_dgrad','conv_wgrad_acc_w','gemm_nn','gemm_tn','linear_wgrad')))
"
  done
done
cp $B htr-vt_b200/libhtrvt_b200.so
