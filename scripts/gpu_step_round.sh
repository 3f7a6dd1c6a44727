#!/bin/bash
# GPU session for the loss head + flat-mode training step.  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/status_step.txt
PH="${PHASES:-loss step train bench}"
for ph in $PH; do
case $ph in
loss)  timeout -k 10 300 python -m pytest tests/test_rpn_loss.py -q -m gpu --timeout 200 > gpurun_out/t_loss.log 2>&1; echo "loss exit $?" >> gpurun_out/status_step.txt ;;
step)  timeout -k 10 400 python -m pytest tests/test_gpu_train_step.py -q -m gpu --timeout 300 > gpurun_out/t_step.log 2>&1; echo "step exit $?" >> gpurun_out/status_step.txt ;;
train) timeout -k 10 600 python -m pytest tests/test_gpu_train.py -q -m gpu --timeout 300 -k "train_graph or train_layer or meta_unit_front or head_out" > gpurun_out/t_train.log 2>&1; echo "train exit $?" >> gpurun_out/status_step.txt ;;
bench) timeout -k 10 400 python scripts/train_bench.py --batch 2 4 > gpurun_out/train_bench2.json 2> gpurun_out/train_bench2.err; echo "bench exit $?" >> gpurun_out/status_step.txt ;;
full)  timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/pytest_full.log 2>&1; echo "full exit $?" >> gpurun_out/status_step.txt ;;
mainbench) timeout -k 10 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "mainbench exit $?" >> gpurun_out/status_step.txt ;;
esac
done
cat gpurun_out/status_step.txt
for f in gpurun_out/t_loss.log gpurun_out/t_step.log gpurun_out/t_train.log; do [ -f $f ] && { echo "== $f"; tail -40 $f; }; done
[ -f gpurun_out/train_bench2.json ] && { cat gpurun_out/train_bench2.json; tail -5 gpurun_out/train_bench2.err; }
