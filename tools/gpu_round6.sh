#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; cat gpurun_out/gpus.txt
echo "== pytest gpu (all) =="; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== dist check N=2 =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; echo "rc=$?"; grep -E "sharded|DIST_CHECK|rror" gpurun_out/dist_check.log | head
echo "== bench N=2 =="; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; grep metric gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
echo "== bench N=1 =="; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench N=1 8192 =="; timeout 900 python bench.py --steps 2 --warmup 3 --size 8192 --no-cpu-baseline > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; echo "rc=$?"; cat gpurun_out/bench_8192.json; tail -3 gpurun_out/bench_8192.err
