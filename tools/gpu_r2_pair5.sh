#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 48 32; do
echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|issuer waits|acc [0-3] quarter" gpurun_out/p.log | tail -20
done
