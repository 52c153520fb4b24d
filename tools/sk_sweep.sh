#!/bin/bash
# sweep the decode kernel's grid size (CTAs = SMs used): exchanges get cheaper with fewer participants
for v in 1.7b 0.6b; do
 for g in 148 128 111 96 74 64 48; do
  echo -n "$v grid=$g: "
  QASR_SK_GRID=$g timeout 120 python tools/decode_ab.py $v 0 2>&1 | tail -1
 done
done
