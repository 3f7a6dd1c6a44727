# A/B the conv kernel variants: parity, per-role cycle profile (RD_CONV_PROF=1) and timing, default vs strip mode.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "conv2d or deconv2d or dla" 2>&1 | tail -3
RD_CONV_STRIP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "conv2d_nhwc" 2>&1 | tail -1
for s in 0 1; do
  RD_CONV_STRIP=$s RD_CONV_PROF=1 timeout 300 python scripts/conv_only.py 2>&1 | grep "rd_conv prof" | awk 'NR==10 || NR==40' 
  echo "STRIP=$s: $(RD_CONV_STRIP=$s timeout 300 python scripts/conv_only.py 2>&1 | tail -1)"
done
