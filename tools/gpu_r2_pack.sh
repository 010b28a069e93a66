#!/bin/bash
# Domain packer with aligned 8-byte gathers (B = 8): accumulator check, parity subset, launch times.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for args in "check 8 256 0 1 0" "check 8 128 0 2 0"; do echo "== probe $args =="; timeout 120 $P $args 2>&1 | grep -E "accumulator|winner check|PROBE|rror"; done
echo "== tests =="; timeout 900 python -m pytest tests -m gpu -q -x -k "parity or large or pair or rgb_tensor" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/pack_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-lena --parity-ranges 0 > /dev/null 2>&1
grep -E "pack_domains|k_umma_refine|k_sort_scatter" gpurun_out/pack_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -8
