#!/bin/bash
RD_BN_STREAM=1 timeout 200 python scripts/bn_small_ab.py
RD_BN_STREAM=0 timeout 200 python scripts/bn_small_ab.py
