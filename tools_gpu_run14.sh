#!/bin/bash
# ncu --set full captures of the kernels added in round 2 (exported to CSV on the box)
set -x
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
NCU="ncu --clock-control none"
exp() { f=gpurun_out/$1; if [ -f $f.ncu-rep ]; then ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null; rm -f $f.ncu-rep; fi; }
T="python bench.py --config TINY --clips 64 --seconds 10 --profile-step --steps 1 --warmup 0"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; shift 4; timeout 600 $NCU --set full -k regex:$regex --launch-skip $skip -c $cnt -f -o gpurun_out/r2o_ncu_$name "$@" > gpurun_out/r2o_ncu_$name.log 2>&1; exp r2o_ncu_$name; tail -2 gpurun_out/r2o_ncu_$name.log | cut -c1-200; }
cap conv0fold conv0_fold_kernel 0 1 $T
cap posconv4 posconv4_kernel 0 1 $T
cap attnblk64 attn_blk_kernel 0 2 $T
cap x3attn "attn_tc_kernel" 0 3 $T --precision bf16x3
cap x3gemm "gemm_tc2_kernel" 8 4 $T --precision bf16x3
cap x3split "split2_rows|split_bf16" 20 3 $T --precision bf16x3
# the two rebuilt wav2vec kernels at the default workload (749-chunk sub-batch)
D="python bench.py --profile-step --steps 1 --warmup 0"
cap conv0fold_full conv0_fold_kernel 0 1 $D
cap posconv4_full posconv4_kernel 0 1 $D
ls -la gpurun_out | grep r2o
echo done
