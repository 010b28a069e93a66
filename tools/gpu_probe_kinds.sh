#!/bin/bash
# Exactness (accumulator dump vs CPU integers) and timing of the tcgen05 search for both MMA kinds.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for kind in 2 1; do
  for args in "check 8 256 0 1 0" "check 8 128 0 3 0" "check 8 128 0 4 0" "check 8 128 0 0 0" "check 4 128 0 4 0" "check 4 128 0 1 0" "check 8 128 0 2 0"; do
    echo "== probe $args kind=$kind =="; timeout 180 $P $args $kind > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -6
  done
done
for kind in 2 1; do
  for d in 0 1 3; do
    echo "== probe time 2048 dbg=$d kind=$kind =="; timeout 300 $P time 8 2048 0 1 $d $kind > gpurun_out/probe_2048_dbg${d}_k$kind.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/probe_2048_dbg${d}_k$kind.log
  done
  echo "== probe time 4096 kind=$kind =="; timeout 600 $P time 8 4096 0 1 0 $kind > gpurun_out/probe_4096_k$kind.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096_k$kind.log
  echo "== probe time 2048 noise kind=$kind =="; timeout 600 $P time 8 2048 0 0 0 $kind > gpurun_out/probe_2048n_k$kind.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_2048n_k$kind.log
  echo "== probe time 2048 B=4 kind=$kind =="; timeout 600 $P time 4 2048 0 1 0 $kind > gpurun_out/probe_b4_2048_k$kind.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b4_2048_k$kind.log
done
