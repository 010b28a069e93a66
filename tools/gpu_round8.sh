#!/bin/bash
mkdir -p gpurun_out
echo "== int8 peak =="; timeout 120 python -c "
import fractal_image_compression_b200 as f
h=f.Handle(0)
for i in range(3): print('int8 peak TOP/s', h.measure_int8_peak())
" > gpurun_out/int8peak.log 2>&1; echo "rc=$?"; cat gpurun_out/int8peak.log
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench =="; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm =="; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
