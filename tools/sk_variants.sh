#!/bin/bash
# A/B of compile-time variants of the streaming decode kernel (rebuilds qasr_stream.o on the box)
run() { for v in 1.7b 0.6b; do echo -n "$1 $v: "; QASR_SK_L2AHEAD=${2:-8} timeout 120 python tools/decode_ab.py $v 0 2>&1 | tail -1; done; }
build() { rm -f smol-vision_b200/csrc/qasr_stream.o; make -C smol-vision_b200/csrc EXTRA="$1" > /dev/null 2>&1 || echo "build failed: $1"; }
run "batch4 maxs4 l2w8"
run "batch4 maxs4 l2w0" 0
run "batch4 maxs4 l2w16" 16
build "-DSK_ATT_BATCH=2"; run "batch2 maxs4 l2w8"
build "-DSK_ATT_BATCH=2 -DSK_ATT_MAXS=8"; run "batch2 maxs8 l2w8"
build "-DSK_ATT_BATCH=1 -DSK_ATT_MAXS=8"; run "batch1 maxs8 l2w8"
build "";
