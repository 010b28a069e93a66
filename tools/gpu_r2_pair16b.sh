#!/bin/bash
# K-split CTA pairs as the default at blockgroesse 16 (grey kind::i8, RGB kind::f16): tests against the oracle, bench lines pair on / off.
mkdir -p gpurun_out
echo "== tests (B=16, rgb tensor, pair, large) =="; timeout 1200 python -m pytest tests -m gpu -q -x -k "b16 or B16 or rgb_tensor or pair or large or tcgen05 or stress or 16" > gpurun_out/pytest_pair16.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_pair16.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-lena"
for a in "--block 16" "--block 16 --pair off" "--rgb --block 16" "--rgb --block 16 --pair off" "--block 16 --pattern noise"; do
  echo "== bench $a =="; timeout 600 $B $a > gpurun_out/b16.json 2> gpurun_out/b16.err; echo "rc=$?"
  python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/b16.json')); r=d['roofline']
    print({k:d[k] for k in ('ms_per_step','value')}, d['config'].get('cta_pairs'), d['config'].get('engine'), 'kernel_ms',r['kernel_ms'],'search_ms',r['search_ms'],'frac',round(r['frac'],3),'bare',round(r.get('frac_of_bare_mma_loop',0),3),'int8nom',round(r.get('frac_of_int8_nominal',0),3), 'parity', d.get('parity_spot'), 'clk', d['clocks']['sm_mhz'])
except Exception as e: print('ERR', e); print(open('gpurun_out/b16.err').read()[-800:])
PY
done
