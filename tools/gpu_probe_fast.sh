#!/bin/bash
# Quick exactness + timing loop for kernel experiments (default kind).
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
K=${1:-0}
for args in "check 8 256 0 1 0" "check 8 128 0 4 0" "check 4 128 0 1 0" "check 8 128 0 2 0"; do
  echo "== probe $args kind=$K =="; timeout 180 $P $args $K > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -6
done
for d in 0 1 3; do
  echo "== probe time 2048 dbg=$d =="; timeout 300 $P time 8 2048 0 1 $d $K > gpurun_out/probe_2048_dbg$d.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/probe_2048_dbg$d.log
done
echo "== probe time 4096 =="; timeout 600 $P time 8 4096 0 1 0 $K > gpurun_out/probe_4096.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096.log
echo "== probe time 2048 B=4 =="; timeout 600 $P time 4 2048 0 1 0 $K > gpurun_out/probe_b4_2048.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b4_2048.log
echo "== probe time 4096 B=16 =="; timeout 600 $P time 16 4096 0 1 0 $K > gpurun_out/probe_b16_4096.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b16_4096.log
