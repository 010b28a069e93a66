#!/bin/bash
# K-split CTA pairs (blockgroesse 16, kind::i8): exactness of every accumulator, winners, timing against the single-CTA kernel.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for args in "check 16 256 32 1 0" "check 16 512 32 2 0" "check 16 256 0 1 0"; do
  echo "== probe $args =="; timeout 120 $P $args > gpurun_out/probe_check16.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch|status" gpurun_out/probe_check16.log | head -8
done
for v in 32 64; do
  for w in 2048 4096; do echo "== probe time 16 $w variant=$v =="; timeout 300 $P time 16 $w $v 1 0 > gpurun_out/probe_t16.log 2>&1; echo "rc=$?"; grep -E "run [12]|umma:|winner|rror" gpurun_out/probe_t16.log; done
done
echo "== probe time 16 4096 noise, pair / single =="; for v in 32 64; do timeout 300 $P time 16 4096 $v 2 0 2>&1 | grep -E "run 2|winner|rror"; done
