#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== chain =="; timeout 300 $P chain | grep -E "free|own commit|nacc=4.*grid=148|nacc=1 .*grid=148" 
for v in 8; do
  for args in "check 8 256 $v 1 0" "check 8 128 $v 4 0" "check 4 128 $v 1 0"; do
    echo "== probe $args =="; timeout 180 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -6
  done
done
for v in 2 8; do
  for d in 0 4; do
    echo "== time 8 2048 variant=$v dbg=$d =="; timeout 300 $P time 8 2048 $v 1 $d > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/p.log
  done
  echo "== time 8 4096 variant=$v =="; timeout 600 $P time 8 4096 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
  echo "== time 4 2048 variant=$v =="; timeout 600 $P time 4 2048 $v 1 0 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run [12]|winner|rror" gpurun_out/p.log
done
echo "== time 8 4096 variant=8 dbg=8 =="; timeout 600 $P time 8 4096 8 1 8 > gpurun_out/p.log 2>&1; echo "rc=$?"; grep -E "run 2|cta 0" gpurun_out/p.log | head -6
echo "== pytest round2+parity =="; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "== lena =="; timeout 600 python bench.py --steps 2 --no-cpu-baseline --parity-ranges 0 2>gpurun_out/b.err | python -c "
import sys,json
d=json.loads(sys.stdin.read())
for r in d['lena']: print(r['case'][:40], r['engine'], 'enc %.1f us dec %.1f us' % (r['gpu_encode_ms']*1e3, r['gpu_decode_ms']*1e3), r['stream_equals_oracle'], r['decode_equals_oracle'])
print('decode', d['decode'])"
