#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== dist check N=8 =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check8.log 2>&1; echo "rc=$?"; grep -E "sharded|DIST_CHECK|rror" gpurun_out/dist_check8.log | head
echo "== bench N=8 =="; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; grep metric gpurun_out/bench_n8.json | cut -c1-900; tail -3 gpurun_out/bench_n8.err
echo "== bench reference N=8 =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_ref_n8.json 2> gpurun_out/bench_ref_n8.err; echo "rc=$?"; grep impl gpurun_out/bench_ref_n8.json | cut -c1-400
