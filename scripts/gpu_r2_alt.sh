#!/bin/bash
# the alternate launch modes stay green: plain stream order, pixel-major kernel for the 128-channel layers, register-staged BN
RD_PDL=0 RD_CONV_T=0 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_parity_full.py -m gpu -q 2>&1 | tail -2
RD_BN_STREAM=0 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py -m gpu -q 2>&1 | tail -2
