#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 48; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0" "check 8 384 $v 3 0"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -8
  done
done
for rep in 1 2; do
  for v in 48 32 64; do
    echo "== time 8 4096 variant=$v (rep $rep) =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  done
done
for d in 1 3 4; do
echo "== time 8 4096 variant=48 dbg=$d =="; timeout 600 $P time 8 4096 48 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|rror" gpurun_out/p.log
done
for v in 48; do
echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|issuer waits|acc [0-3] quarter [01]" gpurun_out/p.log | tail -10
done
echo "== time 8 2048 variant=48 =="; timeout 600 $P time 8 2048 48 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
echo "== time 4 2048 variant=48 =="; timeout 600 $P time 4 2048 48 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
