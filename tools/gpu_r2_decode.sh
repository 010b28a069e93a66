#!/bin/bash
# Decoder sweep over byte-shifted copies: tests, timings of both forms, ncu capture of both sweep kernels.
mkdir -p gpurun_out
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or smoke or facade or golden" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_decode.log
for a in "4096 8" "4096 16" "4096 8 rgb" "8192 8" "256 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -4; done
echo "== ncu: sweep kernels at 4096^2 B=8 =="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_sweep" -s 12 -c 1 -o gpurun_out/r2_k_decode_sweep_sh_4096x4096_B8 python tools/decode_profile.py 4096 8 > gpurun_out/ncu_dec1.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_dec1.log
summ r2_k_decode_sweep_sh_4096x4096_B8 "tools/decode_profile.py 4096 8: a middle sweep of the second decode, shifted copies (k_decode_sweep_sh<1, 8>)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_sweep_v8" -s 12 -c 1 -o gpurun_out/r2_k_decode_sweep_v8_4096x4096_B8 python tools/decode_profile.py 4096 8 > gpurun_out/ncu_dec2.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_dec2.log
summ r2_k_decode_sweep_v8_4096x4096_B8 "tools/decode_profile.py 4096 8: a middle sweep of a decode over one plain plane (k_decode_sweep_v8<1>)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_decode_4096x4096_B8.csv python tools/decode_profile.py 4096 8 > gpurun_out/ncu_dec3.log 2>&1; echo "rc=$?"
rm -f gpurun_out/*.ncu-rep
