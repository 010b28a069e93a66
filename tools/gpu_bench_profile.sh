#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== bench B=8 =="; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench B=16 =="; timeout 900 python bench.py --steps 5 --warmup 3 --block 16 --no-cpu-baseline > gpurun_out/bench_b16.json 2> gpurun_out/bench_b16.err; echo "rc=$?"; cat gpurun_out/bench_b16.json | cut -c1-1600; tail -3 gpurun_out/bench_b16.err
echo "== bench B=4 =="; timeout 900 python bench.py --steps 2 --warmup 3 --block 4 --no-cpu-baseline > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; echo "rc=$?"; cat gpurun_out/bench_b4.json | cut -c1-1600; tail -3 gpurun_out/bench_b4.err
echo "== bench B=8 kind::i8 =="; timeout 900 python bench.py --steps 5 --warmup 3 --mma i8 --no-cpu-baseline > gpurun_out/bench_i8.json 2> gpurun_out/bench_i8.err; echo "rc=$?"; cat gpurun_out/bench_i8.json | cut -c1-1600; tail -3 gpurun_out/bench_i8.err
echo "== bench noise =="; timeout 900 python bench.py --steps 3 --warmup 3 --pattern noise --no-cpu-baseline > gpurun_out/bench_noise.json 2> gpurun_out/bench_noise.err; echo "rc=$?"; cat gpurun_out/bench_noise.json | cut -c1-1600
echo "== ncu launch list (bench) =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full (k_umma_search @4096) =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/prof_umma_r1_f16 $P time 8 4096 0 1 0 2 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
