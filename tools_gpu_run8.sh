#!/bin/bash
# ncu capture of the block-wise AR attention kernel at 256 clips (FULL config, one chunk)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
T="python bench.py --clips 256 --seconds 4 --profile-step --steps 1 --warmup 0"
timeout 900 ncu --clock-control none --set full --import-source on -k regex:attn_blk_kernel --launch-skip 14 -c 4 -f -o gpurun_out/r2i_ncu_attnblk $T > gpurun_out/r2i_ncu_attnblk.log 2>&1
ncu -i gpurun_out/r2i_ncu_attnblk.ncu-rep --page raw --csv > gpurun_out/r2i_ncu_attnblk.raw.csv 2>/dev/null
ncu -i gpurun_out/r2i_ncu_attnblk.ncu-rep --page source --csv > gpurun_out/r2i_ncu_attnblk.source.csv 2>/dev/null
gzip -f gpurun_out/r2i_ncu_attnblk.source.csv
rm -f gpurun_out/r2i_ncu_attnblk.ncu-rep
tail -5 gpurun_out/r2i_ncu_attnblk.log
echo done
