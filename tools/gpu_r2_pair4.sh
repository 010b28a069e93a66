#!/bin/bash
# Two warps per accumulator (variant bit 7) and spinning hand-over waits (bit 8), pair (32) and single (64) kernels.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 160 192; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0" "check 8 384 $v 3 0"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -8
  done
done
for rep in 1 2; do
  for v in 48 160 304 416 64 192; do
    echo "== time 8 4096 variant=$v (rep $rep) =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  done
done
for v in 160 416; do
echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|issuer waits|cta [01] epilogue warp  ?(0|1|4|8|12):" gpurun_out/p.log | tail -12
done
echo "== time 8 4096 noise variant=160 =="; timeout 600 $P time 8 4096 160 2 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
echo "== time 4 2048 variant=160 =="; timeout 600 $P time 4 2048 160 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
echo "== time 4 2048 variant=192 =="; timeout 600 $P time 4 2048 192 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
