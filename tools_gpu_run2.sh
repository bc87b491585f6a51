#!/bin/bash
# GPU session 2 of round 2: tests, launch list of the default bench command, ncu --set full captures per kernel class
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
tail -6 gpurun_out/r2b_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
NCU="ncu --clock-control none"
# launch list of the bench command (whole steps only)
timeout 1500 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2b_launches.csv python bench.py --profile-step --steps 1 --warmup 1 > gpurun_out/r2b_launches.log 2>&1
gzip -f gpurun_out/r2b_launches.csv
# dominant GEMMs of the default workload (layer 0: qkv, out-proj, ffn1, ffn2 after 6 conv layers + projection)
timeout 900 $NCU --set full --import-source on -k regex:gemm_tc2_kernel --launch-skip 7 -c 4 -f -o gpurun_out/r2b_ncu_dominant python bench.py --profile-step --steps 1 --warmup 0 > gpurun_out/r2b_ncu_dominant.log 2>&1
# kernel classes at real shapes on the 2-layer config (same widths, 64 clips x 10 s, FLAME mesh in the step)
T="python bench.py --config TINY --clips 64 --seconds 10 --profile-step --steps 1 --warmup 0"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; timeout 600 $NCU --set full --import-source on -k regex:$regex --launch-skip $skip -c $cnt -f -o gpurun_out/r2b_ncu_$name $T > gpurun_out/r2b_ncu_$name.log 2>&1; }
cap attn attn_tc_kernel 0 12
cap flame flame_tc_kernel 0 1
cap conv0 conv0 0 1
cap ln ln_affine_kernel 40 2
cap adaln adaln_kernel 0 10
cap bits "bits_|bsq_|argmax_bits" 0 12
cap pool "audio_pool|act_cast|savgol|motion_norm|audio_stats" 0 6
cap gemm1 "gemm_tc_kernel" 0 16
cap skinny skinny_gemm_kernel 0 4
timeout 600 $NCU --set full --import-source on -k regex:"resample_mix|split_bf16|vertex_normals|ema_scan" -c 8 -f -o gpurun_out/r2b_ncu_misc python tools_kernels_once.py > gpurun_out/r2b_ncu_misc.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
