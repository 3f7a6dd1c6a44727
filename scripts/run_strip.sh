cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for cfg in "1 1" "1 0" "0 1"; do set -- $cfg
  RD_CONV_STRIP=$1 RD_CONV_BASEOFF=$2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "conv2d_nhwc" 2>&1 | tail -2 > gpurun_out/strip_$1_$2.log
  echo "STRIP=$1 BASEOFF=$2: $(tail -1 gpurun_out/strip_$1_$2.log)"
done
