#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/sorted_fg_time.py | tee gpurun_out/sorted_fg_time.jsonl
