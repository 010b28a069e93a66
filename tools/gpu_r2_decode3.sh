#!/bin/bash
# Decoder sweeps with full-pool codes (every range block points anywhere): warm-L2 ncu captures of both sweep forms.
mkdir -p gpurun_out
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
for a in "4096 8 grey full" "4096 16 grey full"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -3; done
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_decode_sweep_sh" -s 11 -c 1 -o gpurun_out/r2_k_decode_sweep_sh_warm_4096x4096_B8 python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec1.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_dec1.log
summ r2_k_decode_sweep_sh_warm_4096x4096_B8 "tools/decode_profile.py 4096 8 grey full: a middle sweep, shifted copies (k_decode_sweep_sh<1, 8>), full-pool codes, L2 not flushed (--cache-control none)"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"k_decode_sweep_v8" -s 11 -c 1 -o gpurun_out/r2_k_decode_sweep_v8_warm_4096x4096_B8 python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_dec2.log
summ r2_k_decode_sweep_v8_warm_4096x4096_B8 "tools/decode_profile.py 4096 8 grey full: a middle sweep over one plain plane (k_decode_sweep_v8<1>), full-pool codes, L2 not flushed"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/r2_launches_decode_warm_full_4096x4096_B8.csv python tools/decode_profile.py 4096 8 grey full > gpurun_out/ncu_dec3.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_launches_decode_warm_full_4096x4096_B8.csv")) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[ix['Kernel Name']][:50]].append(float(r[ix['Metric Value']]))
for k,v in agg.items(): print("  %-50s n=%3d  min %8.0f  median %8.0f  max %8.0f ns" % (k, len(v), min(v), sorted(v)[len(v)//2], max(v)))
PY
rm -f gpurun_out/*.ncu-rep
