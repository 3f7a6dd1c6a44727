# Conv kernel: parity, per-role cycle profile (RD_CONV_PROF=1), timing, whole-model forward.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -x -k "conv2d or deconv2d or dla" 2>&1 | tail -3
for a in 3 2; do
  RD_CONV_NSA=$a RD_CONV_PROF=1 timeout 300 python scripts/conv_only.py 2>&1 | grep "rd_conv prof" | awk 'NR==40' 
  echo "NSA=$a: $(RD_CONV_NSA=$a timeout 300 python scripts/conv_only.py 2>&1 | tail -1)"
done
timeout 600 python scripts/fwd_bench.py 8 2>&1 | tail -1
RD_MK_PROF=1 timeout 300 python scripts/mk_tc_diag.py 2>&1 | grep "rd_meta prof" | awk '{k=$4; c[k]++; if (c[k]==5) print}'
