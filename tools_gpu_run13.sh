set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python tools_trace.py --clips 256 --seconds 8 --precision bf16x3 --top 60 > gpurun_out/r2n_trace_x3.log 2>&1
head -64 gpurun_out/r2n_trace_x3.log | cut -c1-150
