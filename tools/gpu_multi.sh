#!/bin/bash
# Multi-GPU validation: `gpurun --gpus N -- tools/gpu_multi.sh N` (N = 2, 4, 8).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== multi handle tests =="; timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k multi 2>&1 | tail -4
echo "== C-ABI multi handle, N=$N =="; timeout 900 python tools/multi_abi_bench.py $N > gpurun_out/multi_abi_n$N.json 2> gpurun_out/multi_abi_n$N.err; echo "rc=$?"; cut -c1-2000 gpurun_out/multi_abi_n$N.json; tail -3 gpurun_out/multi_abi_n$N.err
echo "== dist check N=$N =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check$N.log 2>&1; echo "rc=$?"; grep -E "sharded|DIST_CHECK|rror" gpurun_out/dist_check$N.log | head
echo "== bench N=$N =="; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; grep metric gpurun_out/bench_n$N.json | cut -c1-3000; tail -3 gpurun_out/bench_n$N.err
echo "== bench reference N=$N =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "rc=$?"; grep impl gpurun_out/bench_ref_n$N.json | cut -c1-400
