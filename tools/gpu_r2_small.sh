#!/bin/bash
# One-launch decoder for small images (k_decode_small): decode tests, the Lena object of the bench, timings.
mkdir -p gpurun_out
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -x -k "decode or collage or smoke or facade or golden or iso or replay or parity" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_decode.log
for a in "256 8" "512 8" "1024 8 grey full" "256 8 rgb" "512 16 rgb" "2048 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -1 | cut -c1-100; done
echo "== lena bench =="; timeout 600 python tools/lena_bench.py 2>&1 | tail -12
