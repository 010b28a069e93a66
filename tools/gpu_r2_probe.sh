#!/bin/bash
# Round-2 kernel experiments: epilogue mappings of k_umma_search side by side (variant 2 = accumulator per warp,
# 4 = chunk per warp), exactness first, then timings with the probe's debug modes.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== micro =="; timeout 120 $P ldtm; timeout 120 $P mix
for v in 2 4; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0" "check 8 128 $v 2 0" "check 16 256 $v 1 0" "check 8 128 $v 1 0 1"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -6
  done
done
for v in 2 4; do
  for d in 0 1 3 4; do
    echo "== time 8 2048 variant=$v dbg=$d =="; timeout 300 $P time 8 2048 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/p.log
  done
  echo "== time 8 4096 variant=$v =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 8 4096 variant=$v kind=i8 =="; timeout 600 $P time 8 4096 $v 1 0 1 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 16 4096 variant=$v =="; timeout 600 $P time 16 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 4 2048 variant=$v =="; timeout 600 $P time 4 2048 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|cta 0" gpurun_out/p.log | head -18
done
