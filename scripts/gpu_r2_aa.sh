#!/bin/bash
RD_CONV_PROF=1 timeout 120 python scripts/convtc_prof.py 2>&1 | grep -E "==|prof"
