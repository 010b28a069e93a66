#!/bin/bash
# First-contact GPU script: parity tests of the CUDA-core kernels, tcgen05 probe variants
# (each in its own process), then the tcgen05 tests and a short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest direct ==" ; timeout 900 python -m pytest tests -m gpu -x -q -k "not tcgen05 and not large and not engines and not roundtrip_2048 and not smoke" > gpurun_out/pytest_direct.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_direct.log
for v in 0 1; do
  echo "== probe check B=8 W=256 variant $v =="; timeout 120 $P check 8 256 $v 1 > gpurun_out/probe_b8_v$v.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/probe_b8_v$v.log
done
echo "== probe check B=4 W=128 =="; timeout 120 $P check 4 128 0 1 > gpurun_out/probe_b4_v0.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/probe_b4_v0.log
echo "== probe check B=8 W=128 sparse =="; timeout 120 $P check 8 128 0 2 > gpurun_out/probe_b8_sparse.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/probe_b8_sparse.log
echo "== probe time B=8 W=2048 =="; timeout 300 $P time 8 2048 0 1 > gpurun_out/probe_time_2048.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/probe_time_2048.log
echo "== probe time B=8 W=4096 =="; timeout 600 $P time 8 4096 0 1 > gpurun_out/probe_time_4096.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/probe_time_4096.log
echo "== pytest tcgen05 =="; timeout 900 python -m pytest tests -m gpu -x -q -k "tcgen05 or engines or roundtrip_2048" > gpurun_out/pytest_umma.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_umma.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench =="; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
