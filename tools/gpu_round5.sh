#!/bin/bash
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
for args in "check 8 256 0 1 0" "check 4 128 0 1 0" "check 8 128 0 2 0" "check 8 256 0 0 0" "check 4 256 0 2 0"; do
  echo "== probe $args =="; timeout 120 $P $args > gpurun_out/probe_check.log 2>&1; echo "rc=$?"; grep -E "accumulator|winner check|PROBE|rror|mismatch" gpurun_out/probe_check.log | head -8
done
for d in 0 1 3; do
  echo "== probe time 2048 dbg=$d =="; timeout 300 $P time 8 2048 0 1 $d > gpurun_out/probe_2048_dbg$d.log 2>&1; echo "rc=$?"; grep -E "run 2|winner|rror" gpurun_out/probe_2048_dbg$d.log
done
echo "== probe time 4096 =="; timeout 600 $P time 8 4096 0 1 0 > gpurun_out/probe_4096.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096.log
echo "== probe time 4096 noise =="; timeout 600 $P time 8 4096 0 0 0 > gpurun_out/probe_4096_noise.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_4096_noise.log
echo "== probe time 2048 B=4 =="; timeout 600 $P time 4 2048 0 1 0 > gpurun_out/probe_b4_2048.log 2>&1; echo "rc=$?"; grep -E "run 2|umma:|winner|rror" gpurun_out/probe_b4_2048.log
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== bench =="; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list (bench) =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full (k_umma_search @4096) =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/prof_umma_r1f $P time 8 4096 0 1 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
