#!/bin/bash
# Round-2 closing run: every GPU test, smoke, both bench arms, launch list of the bench step, full ncu capture of the
# blockgroesse-16 pair kernel; summaries are made on the box, reports deleted.
mkdir -p gpurun_out
P=fractal-image-compression_b200/lib/umma_probe
summ() { python tools/summarize_ncu.py gpurun_out/$1.ncu-rep gpurun_out/$1 "$2"; rm -f gpurun_out/$1.ncu-rep; }
t0=$(date +%s); echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$? $(( $(date +%s) - t0 )) s"; tail -6 gpurun_out/pytest_gpu.log
echo "== smoke =="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench =="; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err
for a in "--block 16" "--rgb" "--rgb --block 16 --size 4096"; do n=$(echo $a | tr -d ' -'); timeout 600 python bench.py --no-cpu-baseline --no-lena $a > gpurun_out/bench_$n.json 2>/dev/null; echo "bench $a rc=$?"; done
echo "== launch list (bench) =="
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-lena --parity-ranges 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches_bench_final_4096x4096_B8.csv $BENCH > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== full: k_umma_search B=16 pair, 4096 =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_umma_search -c 1 -o gpurun_out/r2_k_umma_search_pair_i8_4096x4096_B16 $P time 16 4096 0 1 0 > gpurun_out/ncu_b16.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_b16.log
summ r2_k_umma_search_pair_i8_4096x4096_B16 "umma_probe time 16 4096 0 1 0 (structured image; default = K-split CTA pairs, kind::i8)"
echo "== full: fused encode + decoder kernels, 4096 (cold L2) =="
timeout 900 ncu --set full --clock-control none -k regex:"k_encode_fused|k_decode_sweep|k_dequant" -c 6 -o gpurun_out/r2_k4_fused_kernels_4096x4096 python tools/k1k4_profile.py > gpurun_out/ncu_k1k4.log 2>&1; echo "rc=$?"
summ r2_k4_fused_kernels_4096x4096 "tools/k1k4_profile.py 4096: fused windowed encode (wk=2), code dequantisation, decoder sweeps (first sweep: implicit start image); L2 flushed before every kernel"
rm -f gpurun_out/*.ncu-rep
