#!/bin/bash
# Round-2 ncu evidence, part 2: captures whose launch offsets skip the once-per-handle f16 self-test (two
# k_umma_search launches of the dump variant) and the first (cold) encode; reports are summarised on the box and deleted.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== sort probe =="; timeout 300 $P sort | grep -E "n=1042441|n=4182025|n=16752649|SORT"
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== bench =="; timeout 900 python bench.py > gpurun_out/bench_r2e.json 2> gpurun_out/b.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print({k:d[k] for k in ('ms_per_step','value','clocks','gpu_launches')}); r=d['roofline']; print(r['kernel_ms'], r['search_ms'], r['frac'], r['frac_of_bare_mma_loop'], r['frac_of_int8_nominal'])
print('e2e',d['e2e']['ms_per_step'],d['e2e']['device_stage_ms']); print('e2e_u8',d['e2e_u8']['ms_per_step'],d['e2e_u8']['device_stage_ms']); print('decode',d['decode']); print('parity',d.get('parity_spot'))
for x in d['lena']: print(x['case'][:40], x['engine'], 'enc %.1f us dec %.1f us' % (x['gpu_encode_ms']*1e3, x['gpu_decode_ms']*1e3), x['stream_equals_oracle'], x['decode_equals_oracle'])
PY
tail -3 gpurun_out/b.err
R8K="python tools/rank_shard_profile.py 1 0"
RANK="python tools/rank_shard_profile.py 8 3"
echo "== full: k_umma_search 8192 (one GPU) =="
$R8K > gpurun_out/plain_8k.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -s 3 -c 1 -o gpurun_out/r2_k_umma_search_f16_8192x8192_B8 $R8K > gpurun_out/ncu_8k.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_8k.log
summ r2_k_umma_search_f16_8192x8192_B8 "tools/rank_shard_profile.py 1 0: the whole 8192^2 search on one GPU (second encode)"
echo "== rank 3 of 8: launch list + search kernel =="
$RANK > gpurun_out/plain_rank.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_rank3of8_8192x8192_B8.csv $RANK > gpurun_out/ncu_rank_l.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_rank.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -s 3 -c 1 -o gpurun_out/r2_k_umma_search_rank3of8_8192x8192_B8 $RANK > gpurun_out/ncu_rank.log 2>&1; echo "rc=$?"
summ r2_k_umma_search_rank3of8_8192x8192_B8 "tools/rank_shard_profile.py 8 3: the search kernel of rank 3 of the 8-GPU run (full pool of 8192^2, 1/8 of the range rows)"
echo "== full: decoder sweep + fused encode =="
python tools/k1k4_profile.py > gpurun_out/plain_k1k4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_encode_fused|k_decode_sweep|k_dequant" -c 6 -o gpurun_out/r2_k4_fused_kernels_4096x4096 python tools/k1k4_profile.py > gpurun_out/ncu_k1k4.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_k1k4.log
summ r2_k4_fused_kernels_4096x4096 "tools/k1k4_profile.py 4096: fused windowed encode (wk=2), code dequantisation, decoder sweeps"
echo "== probe timings of the final kernels =="
for args in "time 8 4096 0 1 0" "time 8 4096 0 0 0" "time 16 4096 0 1 0" "time 4 2048 0 1 0" "time 8 4096 0 1 0 1"; do echo "-- $args"; timeout 300 $P $args 2>&1 | grep -E "run [12]|winner check"; done
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
