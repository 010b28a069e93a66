#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 8 16; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0" "check 8 128 $v 2 0"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror" gpurun_out/probe_check.log | head -4
  done
done
for v in 2 8 16; do
  for d in 0 4; do
    echo "== time 8 2048 variant=$v dbg=$d =="; timeout 300 $P time 8 2048 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|winner check|rror" gpurun_out/p.log
  done
  echo "== time 8 4096 variant=$v =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner check|rror" gpurun_out/p.log
  echo "== time 8 4096 noise variant=$v =="; timeout 600 $P time 8 4096 $v 0 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [2]|winner check|rror" gpurun_out/p.log
  echo "== time 4 2048 variant=$v =="; timeout 600 $P time 4 2048 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [2]|winner check|rror" gpurun_out/p.log
done
echo "== time 8 4096 variant=16 dbg=8 =="; timeout 600 $P time 8 4096 16 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|cta 0" gpurun_out/p.log | head -6
