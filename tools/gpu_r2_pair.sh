#!/bin/bash
# CTA-pair (tcgen05 cta_group::2) form of k_umma_search: bare-loop ceiling, exactness (every accumulator and every
# winner), then timings beside the single-CTA kernel.  Probe variant bit 5 (32) = pair on, bit 6 (64) = pair off.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== peak =="; timeout 120 $P peak 0 0 2>&1 | tail -8
for v in 32 64; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0" "check 8 128 $v 2 0" "check 8 384 $v 3 0"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch|status" gpurun_out/probe_check.log | head -8
  done
done
for v in 32 64; do
  for d in 0 1 3 4; do
    echo "== time 8 2048 variant=$v dbg=$d =="; timeout 300 $P time 8 2048 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/p.log
  done
  echo "== time 8 4096 variant=$v =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  for d in 1 3 4; do
    echo "== time 8 4096 variant=$v dbg=$d =="; timeout 600 $P time 8 4096 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|rror" gpurun_out/p.log
  done
  echo "== time 8 4096 noise variant=$v =="; timeout 600 $P time 8 4096 $v 2 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 4 2048 variant=$v =="; timeout 600 $P time 4 2048 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
done
