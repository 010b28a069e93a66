#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all) =="; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== k1k4 plain =="; timeout 300 python tools/k1k4_profile.py > gpurun_out/k1k4.log 2>&1; echo "rc=$?"; cat gpurun_out/k1k4.log
echo "== bench =="; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu k1k4 =="; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_decimate|k_decode_sweep|k_unpack|k_range_stats|k_domain_stats|k_pack_argb|k_search_direct|k_solve" -c 14 -o gpurun_out/prof_k1k4_r1 python tools/k1k4_profile.py > gpurun_out/ncu_k1k4.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_k1k4.log
