#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bwdsums_ab.py | tee gpurun_out/bwdsums_ab.jsonl
