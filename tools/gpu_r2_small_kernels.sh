#!/bin/bash
# ncu captures of the kernels around the search (refine, domain packer) and of the one-launch small-image decoder.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
timeout 600 ncu --set full --clock-control none -k regex:"k_umma_refine|k_umma_pack_domains|k_umma_pack_ranges" -s 3 -c 3 -o gpurun_out/r2_refine_pack_4096x4096_B8 $P time 8 4096 0 1 0 > gpurun_out/ncu_rp.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_rp.log
summ r2_refine_pack_4096x4096_B8 "umma_probe time 8 4096 0 1 0, second run: k_umma_pack_domains<8, f16>, k_umma_pack_ranges<8, f16>, k_umma_refine<8> (L2 flushed before every kernel)"
timeout 600 ncu --set full --clock-control none -k regex:"k_decode_small" -s 2 -c 1 -o gpurun_out/r2_k_decode_small_256x256_B8 python tools/decode_profile.py 256 8 > gpurun_out/ncu_ds.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_ds.log
summ r2_k_decode_small_256x256_B8 "tools/decode_profile.py 256 8: the one-launch decoder (k_decode_small<1, 8>) on a 256^2 image, 6 sweeps"
rm -f gpurun_out/*.ncu-rep
