#!/bin/bash
# Chunked exact replay of the float avgError sum: unit tests against a sequential float32 sum, 8192^2 decode vs the oracle.
mkdir -p gpurun_out
echo "== replay tests =="; timeout 1200 python -m pytest tests -m gpu -q -x -k "avg_error_replay or decode_8192 or decode_above" > gpurun_out/pytest_replay.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_replay.log
echo "== decode_profile 8192 =="; timeout 300 python tools/decode_profile.py 8192 8 2>&1 | tail -2
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or smoke or facade or golden or iso" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_decode.log
