#!/bin/bash
# bring-up: how often does the 30-step training scenario hit a non-finite gradient, per feature switch
for cfg in "NPVC_OVERLAP=1 NPVC_UMMA_TAP=1" "NPVC_OVERLAP=0 NPVC_UMMA_TAP=1"; do
  bad=0
  for sd in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do
    if env $cfg python tools/train_check.py 3e-4 $sd 2>&1 | grep -q "^step"; then bad=$((bad+1)); fi
  done
  echo "$cfg: $bad / 15 runs with non-finite gradients"
done
