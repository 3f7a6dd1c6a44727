#!/bin/bash
# GPU session for the training-path kernels: each group in its own process under a timeout (a hung
# tcgen05 kernel must not take the other groups down).  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/status_train.txt
PH="${PHASES:-bn dgrad wg_sw wg_noswz}"
for ph in $PH; do
case $ph in
bn)       timeout -k 10 300 python -m pytest tests/test_gpu_train.py -q --timeout 120 -k "bn_act or channel_sums" > gpurun_out/t_bn.log 2>&1; echo "bn exit $?" >> gpurun_out/status_train.txt ;;
dgrad)    timeout -k 10 300 python -m pytest tests/test_gpu_train.py -q --timeout 120 -k "strided_conv_dgrad" > gpurun_out/t_dgrad.log 2>&1; echo "dgrad exit $?" >> gpurun_out/status_train.txt ;;
wg_sw)    timeout -k 10 300 python -m pytest tests/test_gpu_train.py -q --timeout 120 -k "sw128 or is_gradient" > gpurun_out/t_wg_sw.log 2>&1; echo "wg_sw exit $?" >> gpurun_out/status_train.txt ;;
wg_noswz) timeout -k 10 300 python -m pytest tests/test_gpu_train.py -q --timeout 120 -k "noswz" > gpurun_out/t_wg_noswz.log 2>&1; echo "wg_noswz exit $?" >> gpurun_out/status_train.txt ;;
layers)   timeout -k 10 600 python -m pytest tests/test_gpu_train.py -q --timeout 300 -k "train_layer or meta_unit_front or wide_dgrad or head_out or layout" > gpurun_out/t_layers.log 2>&1; echo "layers exit $?" >> gpurun_out/status_train.txt ;;
graph)    timeout -k 10 600 python -m pytest tests/test_gpu_train.py -q --timeout 300 -k "train_graph" > gpurun_out/t_graph.log 2>&1; echo "graph exit $?" >> gpurun_out/status_train.txt ;;
full)     timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/pytest_full.log 2>&1; echo "full exit $?" >> gpurun_out/status_train.txt ;;
trainbench) timeout -k 10 600 python scripts/train_bench.py > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; echo "trainbench exit $?" >> gpurun_out/status_train.txt ;;
esac
done
cat gpurun_out/status_train.txt
for f in gpurun_out/t_*.log; do echo "== $f"; tail -80 $f; done
