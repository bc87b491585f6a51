#!/bin/bash
# GPU session 1 of round 2: sanity of the shared-cudart build, GPU tests, default bench, parity-mode bench lines, FLAME v2 A/B
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
if ! timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; then
  echo "smoke failed with shared cudart; rebuilding static" >> gpurun_out/r2a_smoke.log
  ARTALK_STATIC_CUDART=1 python -m artalk_b200.build --force >> gpurun_out/r2a_smoke.log 2>&1
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2a_smoke.log 2>&1
fi
tail -5 gpurun_out/r2a_smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.json
for p in bf16x3 bf16x6; do
  timeout 900 python bench.py --precision $p --steps 2 --warmup 1 --no-strong --no-cpu-baseline > gpurun_out/r2a_bench_$p.json 2> gpurun_out/r2a_bench_$p.err; echo "bench $p rc=$?"
done
timeout 300 python tools_flame.py > gpurun_out/r2a_flame_v1.json 2>&1
ARTALK_FLAME_V2=1 timeout 300 python tools_flame.py > gpurun_out/r2a_flame_v2.json 2>&1
ARTALK_FLAME_V2=1 timeout 300 python -m pytest tests -m gpu -q -k flame > gpurun_out/r2a_flame_v2_tests.log 2>&1
tail -3 gpurun_out/r2a_flame_v2_tests.log
echo done
