#!/bin/bash
for d in 0 1 3 4; do
  PCGC_UMMA_DBG=$d timeout 60 python tools/bench_conv.py 64 2>&1 | grep -E "dbg=|vrn_a c16|vrn_b c8|deconv_out|vrn_a c32" | sed "s/^/wt=1 /"
done
