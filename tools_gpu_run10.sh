#!/bin/bash
# 2-GPU check of the bench contract (torchrun, NCCL): weak-scaling line with multi_gpu / strong_4096 sub-records
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2k_bench_2gpu.err | cut -c1-300
head -c 3000 gpurun_out/r2k_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2k_ref_2gpu.json 2> gpurun_out/r2k_ref_2gpu.err; echo "ref2 rc=$?"
head -c 600 gpurun_out/r2k_ref_2gpu.json
timeout 300 python -m pytest tests/test_gpu_path.py -m gpu -q -k "non_current_device" --timeout 200 2>&1 | tail -3
echo done
