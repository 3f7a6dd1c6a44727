#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/ab_env.py "AB_BWD_SUMS=0" "AB_BWD_SUMS=1" "AB_BWD_SUMS=auto" | tee gpurun_out/ab_bwdsums2.jsonl
