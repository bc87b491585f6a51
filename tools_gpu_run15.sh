#!/bin/bash
# GPU session 15 of round 2: 256 x 256 single-stage pair tiles for piece-block GEMMs; grouped pinned-host upload
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "split or gemm" --timeout 120 > gpurun_out/r2p_gemm_test.log 2>&1; rc=$?; echo "gemm rc=$rc"
tail -12 gpurun_out/r2p_gemm_test.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "gemm tests failed: stopping"; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2p_pytest.log | cut -c1-400
timeout 900 python tools_ab.py --precision bf16x3 --clips 64 --seconds 10 --rounds 2 --steps 2 base gemm_pair_split=128 gemm_pair_split=0 > gpurun_out/r2p_ab_64x10_x3.json 2> gpurun_out/r2p_ab_64x10_x3.err
cat gpurun_out/r2p_ab_64x10_x3.json; tail -3 gpurun_out/r2p_ab_64x10_x3.err
timeout 900 python bench.py --precision bf16x3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2p_bench_bf16x3.json 2> gpurun_out/r2p_bench_bf16x3.err; echo "bench x3 rc=$?"
head -c 400 gpurun_out/r2p_bench_bf16x3.json; echo
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
echo done
