#!/bin/bash
# Where the pair kernel's accumulator cycle goes: epilogue phase clocks (dbg = 8) for the pipelined variants.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 32 48 72 80; do
echo "== time 8 4096 variant=$v dbg=8 =="; timeout 600 $P time 8 4096 $v 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|issuer waits|cta [01] epilogue warp  ?(0|1|2|3|4|8|12):" gpurun_out/p.log | tail -16
done
