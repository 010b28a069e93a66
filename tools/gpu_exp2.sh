#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for args in "check 8 256 0 1 8" "check 4 128 0 1 8" "check 8 128 0 2 8"; do
  echo "== probe $args =="; timeout 120 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -8
done
for d in 16 8 9 11; do
  echo "== probe time 2048 dbg=$d =="; timeout 300 $P time 8 2048 0 1 $d > gpurun_out/probe_2048_dbg$d.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/probe_2048_dbg$d.log
done
for d in 16 8; do
echo "== probe time 4096 dbg=$d =="; timeout 600 $P time 8 4096 0 1 $d > gpurun_out/probe_4096_$d.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096_$d.log
done
echo "== probe time 4 2048 dbg=8 =="; timeout 600 $P time 4 2048 0 1 8 > gpurun_out/probe_b4_8.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b4_8.log
