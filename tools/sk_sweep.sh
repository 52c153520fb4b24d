#!/bin/bash
# sweep the stall-driven L2 prefetch window (2 KB units per warp) of the streaming decode kernel
for v in 1.7b 0.6b; do
 for a in 0 4 8 12 16 24; do
  echo -n "$v l2window=$a: "
  QASR_SK_L2AHEAD=$a timeout 120 python tools/decode_ab.py $v 0 2>&1 | tail -1
 done
done
