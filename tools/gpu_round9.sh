#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for args in "check 16 256 0 1 0" "check 16 512 0 0 0" "check 8 256 0 1 0"; do
  echo "== probe $args =="; timeout 300 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch|status" gpurun_out/probe_check.log | head -12
done
echo "== probe time 16 2048 =="; timeout 600 $P time 16 2048 0 1 0 > gpurun_out/probe_b16_2048.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b16_2048.log
echo "== probe time 16 4096 =="; timeout 900 $P time 16 4096 0 1 0 > gpurun_out/probe_b16_4096.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b16_4096.log
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu.log
