#!/bin/bash
# Round-2 measurement session of the committed build: GPU suite, smoke, default bench line (bf16, 256 x 30 s) as the driver runs it,
# reference arm, contract-parity line, launch list of the bench command.
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
TAG=${TAG:-r2fin}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --precision bf16x3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16x3.json 2> gpurun_out/${TAG}_bench_bf16x3.err; echo "bf16x3 rc=$?"
timeout 900 python bench.py --precision bf16x6 --steps 2 --warmup 1 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_bench_bf16x6.json 2> gpurun_out/${TAG}_bench_bf16x6.err; echo "bf16x6 rc=$?"
timeout 900 python bench.py --clips 64 --seconds 10 --no-cpu-baseline --no-latency > gpurun_out/${TAG}_bench_64x10.json 2> gpurun_out/${TAG}_bench_64x10.err; echo "64x10 rc=$?"
timeout 1500 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --profile-step --steps 1 --warmup 0 > gpurun_out/${TAG}_launches.log 2>&1
gzip -f gpurun_out/${TAG}_launches.csv
for f in gpurun_out/${TAG}_*.json; do echo "== $f"; head -c 600 $f; echo; done
echo done
