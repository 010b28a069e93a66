#!/bin/bash
# Round-2 batch 2: hand-over chain microbenchmark, the new tests, the full GPU suite, bench with lena / parity objects.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== chain =="; timeout 300 $P chain
echo "== pytest round2 =="; timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -15
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench =="; timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "rc=$?"; cat gpurun_out/bench_r2b.json; tail -5 gpurun_out/bench_r2b.err
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
