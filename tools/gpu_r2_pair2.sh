#!/bin/bash
# Pair vs single kernel, alternating, with the SM clock measured in-kernel (clock64 / globaltimer).
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for rep in 1 2; do
  for v in 32 48 64; do
    echo "== time 8 4096 variant=$v (rep $rep) =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [012]|winner|rror" gpurun_out/p.log
  done
done
for v in 32 48; do
echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|cta [01]" gpurun_out/p.log | head -34
done
for v in 48 32 64; do
  echo "== time 4 2048 variant=$v =="; timeout 600 $P time 4 2048 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
done
nvidia-smi --query-gpu=power.limit,power.draw,clocks.sm,clocks.max.sm,temperature.gpu --format=csv
