#!/bin/bash
# Dev tool: run tools/sweep_dec.py (compress wall, decompress+select wall per decode schedule) under several env settings.
for cfg in "PCGC_ROWS_ON_MAIN=0" "PCGC_ROWS_ON_MAIN=1" "PCGC_SM_LIMIT=132 PCGC_DEC_PAD_KB=24" "PCGC_SM_LIMIT=132 PCGC_DEC_PAD_KB=48" "PCGC_SM_LIMIT=140 PCGC_DEC_PAD_KB=24" "PCGC_SM_LIMIT=124 PCGC_DEC_PAD_KB=24"; do
  echo "== $cfg"
  env $cfg SWEEP_SHORT=1 python tools/sweep_dec.py 2>&1 | tail -5
done
