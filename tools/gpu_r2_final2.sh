#!/bin/bash
# Round-2 closing run, part 2 (one GPU): the 8192^2 workload with the pair kernel (bench line, one rank's share under
# ncu), decoder timings after the 32-bit strip arithmetic.
mkdir -p gpurun_out
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== decode tests =="; timeout 900 python -m pytest tests -m gpu -q -k "decode or collage or golden" 2>&1 | tail -2
for a in "4096 8 grey full" "4096 16 grey full" "4096 8 rgb" "8192 8" "256 8"; do echo "== decode_profile $a =="; timeout 300 python tools/decode_profile.py $a 2>&1 | tail -1; done
echo "== bench --size 8192 (one GPU) =="; timeout 900 python bench.py --size 8192 --steps 2 --warmup 3 --no-cpu-baseline --no-lena > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; echo "rc=$?"; tail -2 gpurun_out/bench_8192.err
RANK="python tools/rank_shard_profile.py 8 3"
echo "== rank 3 of 8: launch list + search kernel (pairs) =="
$RANK > gpurun_out/plain_rank.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_rank3of8_pair_8192x8192_B8.csv $RANK > gpurun_out/ncu_rank_l.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_rank.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -s 3 -c 1 -o gpurun_out/r2_k_umma_search_pair_rank3of8_8192x8192_B8 $RANK > gpurun_out/ncu_rank.log 2>&1; echo "rc=$?"
summ r2_k_umma_search_pair_rank3of8_8192x8192_B8 "tools/rank_shard_profile.py 8 3: the search kernel rank 3 of the 8-GPU run executes (CTA pairs; full pool of 8192^2, 1/8 of the range rows), second encode"
rm -f gpurun_out/*.ncu-rep
