#!/bin/bash
# Decoder: implicit start image (first sweep reads nothing), 32-bit domain loads at B = 16 -- tests and timings.
mkdir -p gpurun_out
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or smoke or facade or golden or iso" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_decode.log
for a in "4096 8 grey full" "4096 16 grey full" "4096 8 rgb" "256 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -2; done
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import bench
PY
