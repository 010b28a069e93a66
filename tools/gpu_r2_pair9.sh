#!/bin/bash
# GPU tests of the pair kernel + ncu evidence (launch list of the bench step, one full capture at 4096^2, one of rank 3 of 8).
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0"
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== launch list (bench) =="
$BENCH > gpurun_out/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench_pair_4096x4096_B8.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: k_umma_search (pair) 4096 =="
$P time 8 4096 0 1 0 > gpurun_out/plain_probe.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/r2_k_umma_search_pair_f16_4096x4096_B8 $P time 8 4096 0 1 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_full.log; grep -E "run [12]|winner" gpurun_out/plain_probe.log
summ r2_k_umma_search_pair_f16_4096x4096_B8 "umma_probe time 8 4096 0 1 0 (structured image; default = CTA pairs, deferred-test epilogue)"
echo "== full: one rank of N=8 at 8192 =="
python tools/rank_shard_profile.py 8 3 > gpurun_out/plain_rank.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_umma_search" -s 1 -c 1 -o gpurun_out/r2_k_umma_search_pair_rank3of8_8192x8192_B8 python tools/rank_shard_profile.py 8 3 > gpurun_out/ncu_rank.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_rank.log; tail -1 gpurun_out/ncu_rank.log
summ r2_k_umma_search_pair_rank3of8_8192x8192_B8 "tools/rank_shard_profile.py 8 3: the search kernel rank 3 of the 8-GPU run executes (CTA pairs), second encode"
rm -f gpurun_out/*.ncu-rep
