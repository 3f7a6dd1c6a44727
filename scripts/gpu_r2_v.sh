#!/bin/bash
RD_CONVT_PROF=1 timeout 120 python scripts/convt_prof.py 2>&1 | grep -E "==|prof" 
