set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "pos_conv_four or conv0 or block_wise or fp32_grade or split_gemm" --timeout 1200 > gpurun_out/r2w_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -15 gpurun_out/r2w_memcheck.log | cut -c1-300
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2w_memcheck.log
