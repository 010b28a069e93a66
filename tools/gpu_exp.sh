#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for d in 32 16 17; do
  echo "== probe time 2048 dbg=$d =="; timeout 300 $P time 8 2048 0 1 $d > gpurun_out/probe_2048_dbg$d.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/probe_2048_dbg$d.log
done
echo "== probe time 4096 dbg=32 =="; timeout 600 $P time 8 4096 0 1 32 > gpurun_out/probe_4096_32.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096_32.log
