#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for v in 2 8; do
  for d in 0 1 3 4; do
    echo "== time 8 4096 variant=$v dbg=$d =="; timeout 300 $P time 8 4096 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|rror" gpurun_out/p.log
  done
done
echo "== time 8 8192 default =="; timeout 600 $P time 8 8192 0 1 3 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|rror" gpurun_out/p.log
echo "== peak =="; timeout 120 $P peak 8 128
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench =="; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print({k:d[k] for k in ('ms_per_step','value','clocks','gpu_launches')})
print(d['roofline'])
print('e2e',d['e2e']); print('e2e_u8',d.get('e2e_u8')); print('decode',d['decode']); print('parity',d.get('parity_spot'))
for r in d['lena']: print(r['case'][:40], r['engine'], 'enc %.1f us dec %.1f us' % (r['gpu_encode_ms']*1e3, r['gpu_decode_ms']*1e3), r['stream_equals_oracle'], r['decode_equals_oracle'])
PY
tail -3 gpurun_out/bench_r2c.err
