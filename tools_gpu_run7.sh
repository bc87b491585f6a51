set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python tools_trace.py --clips 256 --seconds 8 --top 70 > gpurun_out/r2h_trace_blk1.log 2>&1
ARTALK_ATTN_BLK=0 timeout 600 python tools_trace.py --clips 256 --seconds 8 --top 70 > gpurun_out/r2h_trace_blk0.log 2>&1
grep -i "attention" gpurun_out/r2h_trace_blk1.log | head -20
grep -i "attention" gpurun_out/r2h_trace_blk0.log | head -20
