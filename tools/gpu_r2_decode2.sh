#!/bin/bash
# Decoder sweep over byte-shifted copies, second pass: tests, timings, launch lists with warm and cold L2.
mkdir -p gpurun_out
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or smoke or facade or golden" > gpurun_out/pytest_decode.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_decode.log
for a in "4096 8" "4096 16" "4096 8 rgb" "2048 8" "256 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -3; done
echo "== launch list, warm L2 (--cache-control none) =="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/r2_launches_decode_warm_4096x4096_B8.csv python tools/decode_profile.py 4096 8 > gpurun_out/ncu_dec3.log 2>&1; echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r2_launches_decode_warm_4096x4096_B8_rgb.csv python tools/decode_profile.py 4096 8 rgb > gpurun_out/ncu_dec4.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv, collections
for f in ("r2_launches_decode_warm_4096x4096_B8.csv", "r2_launches_decode_warm_4096x4096_B8_rgb.csv"):
    rows=[r for r in csv.reader(open("gpurun_out/"+f)) if len(r)>10]
    ix={h:i for i,h in enumerate(rows[0])}
    agg=collections.defaultdict(list)
    for r in rows[1:]:
        agg[r[ix['Kernel Name']][:50]].append(float(r[ix['Metric Value']]))
    print(f)
    for k,v in agg.items(): print("  %-50s n=%3d  min %8.0f  median %8.0f  max %8.0f ns" % (k, len(v), min(v), sorted(v)[len(v)//2], max(v)))
PY
