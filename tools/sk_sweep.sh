#!/bin/bash
# sweep the stall-driven L2 prefetch: window (2 KB units per warp beyond the ring) x units issued per poll iteration
for v in 1.7b 0.6b; do
 for a in "8 2" "8 3" "8 4" "12 3" "16 4" "16 8"; do
  set -- $a
  echo -n "$v l2window=$1 issue=$2: "
  QASR_SK_L2AHEAD=$1 QASR_SK_L2ISSUE=$2 timeout 120 python tools/decode_ab.py $v 0 2>&1 | tail -1
 done
done
