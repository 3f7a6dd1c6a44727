#!/bin/bash
# One GPU-box session: parity tests in isolated phases (a hung tcgen05 kernel must not take the
# other phases down), smoke, bench, ncu launch list + full capture.  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PH="${PHASES:-base tc meta smoke bench ncu}"
for ph in $PH; do
case $ph in
base)  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "not tc_probe and not tma_probe and not meta_kernel and not conv2d and not deconv2d and not dla_backbone" > gpurun_out/pytest_base.log 2>&1; echo "base exit $?" >> gpurun_out/status.txt ;;
tma)   RD_TMA_PROBE_VERBOSE=1 timeout 300 compute-sanitizer python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "tma_probe" -s > gpurun_out/pytest_tma.log 2>&1; echo "tma exit $?" >> gpurun_out/status.txt ;;
dbgws) timeout 300 python scripts/dbg_ws.py fwd > gpurun_out/dbg_ws_fwd.log 2>&1; echo "dbgws-fwd exit $?" >> gpurun_out/status.txt
       timeout 300 python scripts/dbg_ws.py bwd > gpurun_out/dbg_ws_bwd.log 2>&1; echo "dbgws-bwd exit $?" >> gpurun_out/status.txt ;;
tc)    timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 120 -k "tc_probe" > gpurun_out/pytest_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/status.txt ;;
meta)  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "meta_kernel" > gpurun_out/pytest_meta.log 2>&1; echo "meta exit $?" >> gpurun_out/status.txt ;;
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/status.txt ;;
bench) timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?" >> gpurun_out/status.txt ;;
ncu)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/status.txt
       timeout 900 ncu --set full --clock-control none --import-source on -k regex:meta_ -s 8 -c 4 -o gpurun_out/prof_meta -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/status.txt ;;
conv)  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "conv2d or deconv2d" > gpurun_out/pytest_conv.log 2>&1; echo "conv exit $?" >> gpurun_out/status.txt ;;
ncuconv) timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_fprop -s 4 -c 3 -o gpurun_out/prof_conv -f python scripts/mk_tc_diag.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu-conv exit $?" >> gpurun_out/status.txt ;;
model) timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "dla_backbone" > gpurun_out/pytest_model.log 2>&1; echo "model exit $?" >> gpurun_out/status.txt
       timeout 600 python scripts/fwd_bench.py 8 > gpurun_out/fwd_bench.json 2> gpurun_out/fwd_bench.err; echo "fwdbench exit $?" >> gpurun_out/status.txt ;;
full)  timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/pytest_full.log 2>&1; echo "full exit $?" >> gpurun_out/status.txt ;;
ncuall) timeout 900 ncu --set full --clock-control none -k regex:"meta_ws|conv_kernel" -s 20 -c 12 -o gpurun_out/prof_all -f python scripts/mk_tc_diag.py > gpurun_out/ncu_all.log 2>&1; echo "ncu-all exit $?" >> gpurun_out/status.txt ;;
diag)  timeout 600 python scripts/mk_tc_diag.py > gpurun_out/mk_tc_diag.json 2> gpurun_out/mk_tc_diag.err; echo "diag exit $?" >> gpurun_out/status.txt ;;
ncutc) timeout 900 ncu --set full --clock-control none --import-source on -k regex:meta_fwd_tc -s 3 -c 1 -o gpurun_out/prof_meta_tc -f python scripts/mk_tc_diag.py 0 > gpurun_out/ncu_tc.log 2>&1; echo "ncu-tc exit $?" >> gpurun_out/status.txt ;;
ncuws) timeout 900 ncu --set full --clock-control none --import-source on -k regex:meta_ws -s 3 -c 5 -o gpurun_out/prof_meta_ws -f python scripts/mk_tc_diag.py > gpurun_out/ncu_ws.log 2>&1; echo "ncu-ws exit $?" >> gpurun_out/status.txt ;;
esac
done
cat gpurun_out/status.txt
tail -5 gpurun_out/pytest_base.log gpurun_out/pytest_tc.log gpurun_out/pytest_meta.log 2>/dev/null
