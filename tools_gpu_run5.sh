#!/bin/bash
# GPU session 5 of round 2: conv layer 0 with the LayerNorm folded through the conv (conv0_fold.cu): op test, full suite, A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv0" --timeout 200 > gpurun_out/r2e_conv0_test.log 2>&1; echo "conv0 rc=$?"
tail -25 gpurun_out/r2e_conv0_test.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2e_pytest.log
timeout 600 python tools_ab.py --clips 64 --seconds 10 --rounds 3 base conv0_fold=0 > gpurun_out/r2e_ab_64x10.json 2> gpurun_out/r2e_ab_64x10.err
cat gpurun_out/r2e_ab_64x10.json; tail -3 gpurun_out/r2e_ab_64x10.err
timeout 900 python tools_ab.py --clips 256 --seconds 30 --rounds 2 --steps 2 base conv0_fold=0 > gpurun_out/r2e_ab_256x30.json 2> gpurun_out/r2e_ab_256x30.err
cat gpurun_out/r2e_ab_256x30.json; tail -3 gpurun_out/r2e_ab_256x30.err
echo done
