#!/bin/bash
# GPU session 11 of round 2: parity-grade attention on the tensor cores (SPLIT variant of attn_tc_kernel)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" --timeout 120 > gpurun_out/r2l_attn_test.log 2>&1; rc=$?; echo "attn rc=$rc"
tail -30 gpurun_out/r2l_attn_test.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "attention tests failed: stopping"; exit 0; fi
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2l_pytest.log | cut -c1-400
grep bf16x3 gpurun_out/parity_measured.jsonl | cut -c1-400
timeout 900 python tools_ab.py --precision bf16x3 --clips 64 --seconds 10 --rounds 2 --steps 2 base attn_split=0 > gpurun_out/r2l_ab_64x10_x3.json 2> gpurun_out/r2l_ab_64x10_x3.err
cat gpurun_out/r2l_ab_64x10_x3.json; tail -3 gpurun_out/r2l_ab_64x10_x3.err
timeout 900 python bench.py --precision bf16x3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2l_bench_bf16x3.json 2> gpurun_out/r2l_bench_bf16x3.err; echo "bench x3 rc=$?"
head -c 400 gpurun_out/r2l_bench_bf16x3.json; echo
echo done
