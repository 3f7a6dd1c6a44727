#!/bin/bash
# End-of-round GPU session: full -m gpu suite, smoke, default bench, in-situ breakdowns.  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/status_final.txt
timeout -k 10 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 > gpurun_out/pytest_full.log 2>&1; echo "full exit $?" >> gpurun_out/status_final.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/status_final.txt
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?" >> gpurun_out/status_final.txt
timeout 300 python scripts/train_profile.py 4 > gpurun_out/train_insitu_b4.txt 2>&1; echo "insitu exit $?" >> gpurun_out/status_final.txt
timeout 300 python scripts/conv_shapes_profile.py 4 > gpurun_out/conv_shapes_b4.txt 2>&1; echo "shapes exit $?" >> gpurun_out/status_final.txt
timeout 300 python scripts/bn_bench.py 64 128 576 > gpurun_out/bn_bench.json 2>&1; echo "bnbench exit $?" >> gpurun_out/status_final.txt
cat gpurun_out/status_final.txt; tail -4 gpurun_out/pytest_full.log; tail -3 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench_default.json
