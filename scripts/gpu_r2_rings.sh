#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/ab_env.py "" "RD_CONVT_RINGS=24" "RD_CONVT_RINGS=33" | tee gpurun_out/ab_rings.jsonl
