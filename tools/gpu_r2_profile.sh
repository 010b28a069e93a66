#!/bin/bash
# Round-2 ncu evidence (one gpurun call; every ncu run follows a plain run of the same command that exited 0).
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0"
echo "== launch list (bench) =="
$BENCH > gpurun_out/plain_bench.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: k_umma_search 4096 =="
$P time 8 4096 0 1 0 > gpurun_out/plain_probe.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/r2_prof_search_4096 $P time 8 4096 0 1 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full.log
echo "== full: k_umma_search 8192 =="
$P time 8 8192 0 1 0 > gpurun_out/plain_probe8k.log 2>&1 && timeout 1200 ncu --set full --clock-control none -k regex:k_umma_search -c 1 -o gpurun_out/r2_prof_search_8192 $P time 8 8192 0 1 0 > gpurun_out/ncu_full8k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_full8k.log
echo "== full: one rank of N=8 at 8192 =="
python tools/rank_shard_profile.py 8 3 > gpurun_out/plain_rank.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:"k_umma_search|k_umma_pack|k_umma_refine|DeviceRadixSort" -s 12 -c 12 -o gpurun_out/r2_prof_rank3of8_8192 python tools/rank_shard_profile.py 8 3 > gpurun_out/ncu_rank.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_rank.log; tail -2 gpurun_out/ncu_rank.log
echo "== full: K1 / K4 / fused =="
python tools/k1k4_profile.py > gpurun_out/plain_k1k4.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:"k_encode_fused|k_decode_sweep|k_decimate|k_domain_stats|k_range_stats|k_sweep_finish|k_dequant" -c 24 -o gpurun_out/r2_prof_k1k4 python tools/k1k4_profile.py > gpurun_out/ncu_k1k4.log 2>&1; echo "rc=$?"; cat gpurun_out/plain_k1k4.log
ls -la gpurun_out/*.ncu-rep
