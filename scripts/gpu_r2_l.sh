#!/bin/bash
for cfg in "220 0" "100 0" "100 16384" "70 16384" "100 12288" "220 16384" "220 0"; do
  set -- $cfg
  if [ "$2" = "0" ]; then RD_WGRAD_SMEM_KB=$1 timeout 200 python scripts/ab_overlap.py 2>&1 | tail -1
  else RD_WGRAD_SMEM_KB=$1 RD_BN_UB=$2 timeout 200 python scripts/ab_overlap.py 2>&1 | tail -1; fi
done
