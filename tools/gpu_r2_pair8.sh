#!/bin/bash
# Full GPU test suite + smoke + bench lines of the CTA-pair kernel (default) and of the single-CTA kernel.
mkdir -p gpurun_out
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench (default: pairs) =="; timeout 900 python bench.py > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "rc=$?"; cat gpurun_out/bench_pair.json; tail -3 gpurun_out/bench_pair.err
echo "== bench --pair off =="; timeout 900 python bench.py --pair off --no-cpu-baseline --no-lena > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "rc=$?"; cat gpurun_out/bench_single.json; tail -3 gpurun_out/bench_single.err
echo "== bench rgb =="; timeout 900 python bench.py --rgb --no-cpu-baseline --no-lena > gpurun_out/bench_rgb_pair.json 2> gpurun_out/bench_rgb_pair.err; echo "rc=$?"; cat gpurun_out/bench_rgb_pair.json; tail -3 gpurun_out/bench_rgb_pair.err
