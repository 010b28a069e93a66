#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for d in 0 4 1 3; do
  echo "== probe time 2048 dbg=$d =="; timeout 300 $P time 8 2048 0 1 $d > gpurun_out/probe_2048_dbg$d.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:" gpurun_out/probe_2048_dbg$d.log
done
echo "== probe time 2048 noise =="; timeout 300 $P time 8 2048 0 0 0 > gpurun_out/probe_2048_noise.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner" gpurun_out/probe_2048_noise.log
echo "== probe time 4096 B=4 W=2048 =="; timeout 300 $P time 4 2048 0 1 0 > gpurun_out/probe_b4_2048.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner" gpurun_out/probe_b4_2048.log
echo "== bench =="; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench reference arm =="; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== ncu launch list (bench) =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_bench.log
echo "== ncu full (k_umma_search @2048) =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/prof_umma_r1 $P time 8 2048 0 1 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
