# A/B the conv kernel variants: per-role cycle profile (RD_CONV_PROF=1) and timing, default vs strip mode.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for s in 0 1; do
  RD_CONV_STRIP=$s RD_CONV_PROF=1 timeout 300 python scripts/conv_only.py 2>&1 | grep "rd_conv prof" | awk 'NR==10 || NR==40' 
  echo "STRIP=$s: $(RD_CONV_STRIP=$s timeout 300 python scripts/conv_only.py 2>&1 | tail -1)"
done
