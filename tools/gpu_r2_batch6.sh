#!/bin/bash
mkdir -p gpurun_out
echo "== pytest rgb tensor =="; timeout 900 python -m pytest tests/test_gpu_rgb_tensor.py -m gpu -x -q 2>&1 | tail -15
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== bench rgb b16 2048 =="; timeout 600 python bench.py --rgb --block 16 --size 2048 --steps 3 --no-cpu-baseline --no-lena > gpurun_out/bench_rgb_b16_2048.json 2> gpurun_out/b.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_rgb_b16_2048.json'))
print({k:d[k] for k in ('ms_per_step','value','config')}); print(d['roofline']); print(d.get('parity_spot')); print(d['decode'])
PY
tail -3 gpurun_out/b.err
echo "== bench rgb b16 4096 =="; timeout 600 python bench.py --rgb --block 16 --size 4096 --steps 3 --no-cpu-baseline --no-lena > gpurun_out/bench_rgb_b16_4096.json 2> gpurun_out/b.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_rgb_b16_4096.json'))
print({k:d[k] for k in ('ms_per_step','value')}); print(d['roofline']); print(d.get('parity_spot'))
PY
tail -3 gpurun_out/b.err
echo "== bench default =="; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/b.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2d.json'))
print({k:d[k] for k in ('ms_per_step','value','clocks')}); print(d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_of_bare_mma_loop'])
print('e2e',d['e2e']); print('e2e_u8',d.get('e2e_u8')); print('decode',d['decode']); print('parity',d.get('parity_spot'))
for r in d['lena']: print(r['case'][:40], r['engine'], 'enc %.1f us dec %.1f us' % (r['gpu_encode_ms']*1e3, r['gpu_decode_ms']*1e3), r['stream_equals_oracle'], r['decode_equals_oracle'])
PY
tail -3 gpurun_out/b.err
