#!/bin/bash
# Round-end rehearsal: what the driver runs (GPU tests, smoke, both bench arms), with wall-clock per leg.
mkdir -p gpurun_out
t0=$(date +%s)
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$? $(( $(date +%s) - t0 )) s"; tail -6 gpurun_out/pytest_gpu.log
t0=$(date +%s)
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/smoke.log
t0=$(date +%s)
echo "== bench reference arm =="; timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$? $(( $(date +%s) - t0 )) s"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
t0=$(date +%s)
echo "== bench =="; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$? $(( $(date +%s) - t0 )) s"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
