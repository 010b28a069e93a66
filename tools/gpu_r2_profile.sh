#!/bin/bash
# Round-2 ncu evidence (one gpurun call; every ncu run follows a plain run of the same command that exited 0).
# The reports are summarised ON the box (tools/summarize_ncu.py) and deleted: gpurun_out/ carries at most 64 MiB back.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0"
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
echo "== sort probe =="; timeout 300 $P sort | tail -30
R8K="python tools/rank_shard_profile.py 1 0"
echo "== launch list (bench) =="
$BENCH > gpurun_out/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench_4096x4096_B8.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: k_umma_search 4096 =="
$P time 8 4096 0 1 0 > gpurun_out/plain_probe.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/r2_k_umma_search_f16_4096x4096_B8 $P time 8 4096 0 1 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_full.log
summ r2_k_umma_search_f16_4096x4096_B8 "umma_probe time 8 4096 0 1 0 (structured image, default epilogue variant)"
echo "== full: k_umma_search 8192 =="
$R8K > gpurun_out/plain_probe8k.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -s 1 -c 1 -o gpurun_out/r2_k_umma_search_f16_8192x8192_B8 $R8K > gpurun_out/ncu_full8k.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_full8k.log; cat gpurun_out/plain_probe8k.log
summ r2_k_umma_search_f16_8192x8192_B8 "tools/rank_shard_profile.py 1 0 (the whole 8192^2 search on one GPU, second encode)"
echo "== full: one rank of N=8 at 8192 =="
python tools/rank_shard_profile.py 8 3 > gpurun_out/plain_rank.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_umma_search|k_umma_pack|k_umma_refine|k_sort|k_umma_sortkeys|k_domain_stats|k_decimate|k_range_stats|k_solve" -s 18 -c 18 -o gpurun_out/r2_rank3of8_8192x8192_B8 python tools/rank_shard_profile.py 8 3 > gpurun_out/ncu_rank.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_rank.log; tail -1 gpurun_out/ncu_rank.log
summ r2_rank3of8_8192x8192_B8 "tools/rank_shard_profile.py 8 3: what rank 3 of the 8-GPU run executes (full pool of 8192^2, 1/8 of the range rows), second encode"
echo "== full: K1 / K4 / fused =="
python tools/k1k4_profile.py > gpurun_out/plain_k1k4.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_encode_fused|k_decode_sweep|k_decimate|k_domain_stats|k_range_stats|k_sweep_finish|k_dequant|k_search_direct" -c 20 -o gpurun_out/r2_k1_k4_kernels_4096x4096 python tools/k1k4_profile.py > gpurun_out/ncu_k1k4.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_k1k4.log
summ r2_k1_k4_kernels_4096x4096 "tools/k1k4_profile.py 4096: fused windowed encode, K1 + direct search, decoder sweeps"
echo "== full: RGB B=16 =="
python bench.py --rgb --block 16 --size 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0 > gpurun_out/plain_rgb16.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_umma_search" -s 3 -c 1 -o gpurun_out/r2_k_umma_search_rgb_2048x2048_B16 python bench.py --rgb --block 16 --size 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0 > gpurun_out/ncu_rgb16.log 2>&1; echo "rc=$?"
summ r2_k_umma_search_rgb_2048x2048_B16 "bench.py --rgb --block 16 --size 2048: the K-split binary16 search (256-row super-blocks)"
rm -f gpurun_out/*.ncu-rep; ls -la gpurun_out/ | head -40; du -sh gpurun_out
