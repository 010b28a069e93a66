#!/bin/bash
# Decoder loop over the row-pair interleaved decimated plane: tests, timings, warm ncu capture.
mkdir -p gpurun_out
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or smoke or facade or golden or iso or replay" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_decode.log
for a in "4096 8 grey full" "4096 16 grey full" "4096 8 rgb" "8192 8" "256 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -1; done
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_decode_sweep_il" -s 11 -c 1 -o gpurun_out/r2_k_decode_sweep_il_warm_4096x4096_B8 python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_dec1.log
summ r2_k_decode_sweep_il_warm_4096x4096_B8 "tools/decode_profile.py 4096 8 grey full: a middle sweep over the row-pair interleaved plane (k_decode_sweep_il<1, 8>), full-pool codes, L2 not flushed (--cache-control none)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_sweep_il" -s 11 -c 1 -o gpurun_out/r2_k_decode_sweep_il_cold_4096x4096_B8 python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec2.log 2>&1; echo "rc=$?"
summ r2_k_decode_sweep_il_cold_4096x4096_B8 "tools/decode_profile.py 4096 8 grey full: the same sweep with the L2 flushed before the kernel (ncu default)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 60 --csv --log-file gpurun_out/r2_launches_decode_il_warm_4096x4096_B8.csv python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec3.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_decode_il_warm_4096x4096_B8.csv")) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[ix['Kernel Name']][:60]].append(float(r[ix['Metric Value']]))
for k,v in agg.items():
    if 'decode' in k or 'dequant' in k: print("  %-60s n=%3d  min %8.0f  median %8.0f  max %8.0f ns" % (k, len(v), min(v), sorted(v)[len(v)//2], max(v)))
PY
rm -f gpurun_out/*.ncu-rep
