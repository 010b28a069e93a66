#!/bin/bash
# Multi-GPU lines only: `gpurun --gpus N -- tools/gpu_multi_short.sh N` (C-ABI multi handle + torchrun bench).
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== C-ABI multi handle, N=$N =="; timeout 600 python tools/multi_abi_bench.py $N > gpurun_out/multi_abi_n$N.json 2> gpurun_out/multi_abi_n$N.err; echo "rc=$?"; cut -c1-1200 gpurun_out/multi_abi_n$N.json; tail -2 gpurun_out/multi_abi_n$N.err
echo "== bench N=$N =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; grep metric gpurun_out/bench_n$N.json | cut -c1-1500; tail -2 gpurun_out/bench_n$N.err
