#!/bin/bash
# GPU session 17 of round 2: single-pass (lazy maximum) softmax for the unbounded ping-pong attention items
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" --timeout 120 > gpurun_out/r2s_attn_test.log 2>&1; rc=$?; echo "attn rc=$rc"
tail -25 gpurun_out/r2s_attn_test.log | cut -c1-300
if [ $rc -ne 0 ]; then echo "attention tests failed: stopping"; exit 0; fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2s_pytest.log | cut -c1-400
timeout 600 python tools_ab.py --clips 64 --seconds 10 --rounds 3 base attn_lazy=0 > gpurun_out/r2s_ab_64x10.json 2> gpurun_out/r2s_ab_64x10.err
cat gpurun_out/r2s_ab_64x10.json; tail -3 gpurun_out/r2s_ab_64x10.err
timeout 900 python tools_ab.py --clips 256 --seconds 30 --rounds 2 --steps 2 base attn_lazy=0 > gpurun_out/r2s_ab_256x30.json 2> gpurun_out/r2s_ab_256x30.err
cat gpurun_out/r2s_ab_256x30.json; tail -3 gpurun_out/r2s_ab_256x30.err
echo done
