#!/bin/bash
# GPU session 3 of round 2: tests, option A/Bs, parity-mode bench lines, ncu captures exported to CSV on the box (reports are
# too large to copy back), launch list of the default bench command
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
rm -f gpurun_out/parity_measured.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 600 python tools_ab.py --clips 64 --seconds 10 --rounds 3 base attn_bound=0 attn_poly=1 attn_poly=2 > gpurun_out/r2c_ab_64x10.json 2> gpurun_out/r2c_ab_64x10.err
cat gpurun_out/r2c_ab_64x10.json
timeout 900 python tools_ab.py --clips 256 --seconds 30 --rounds 2 --steps 2 base attn_bound=0 attn_poly=1 attn_poly=2 > gpurun_out/r2c_ab_256x30.json 2> gpurun_out/r2c_ab_256x30.err
cat gpurun_out/r2c_ab_256x30.json
timeout 900 python bench.py --steps 3 --warmup 2 --no-strong --no-cpu-baseline > gpurun_out/r2c_bench_bf16.json 2> gpurun_out/r2c_bench_bf16.err; echo "bench bf16 rc=$?"
for p in bf16x3 bf16x6; do
  timeout 900 python bench.py --precision $p --steps 2 --warmup 1 --no-strong --no-cpu-baseline > gpurun_out/r2c_bench_$p.json 2> gpurun_out/r2c_bench_$p.err; echo "bench $p rc=$?"
done
NCU="ncu --clock-control none"
exp() { f=gpurun_out/$1; if [ -f $f.ncu-rep ]; then ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null; if [ -n "$2" ]; then ncu -i $f.ncu-rep --page source --csv > $f.source.csv 2>/dev/null; gzip -f $f.source.csv; fi; rm -f $f.ncu-rep; fi; }
# dominant GEMMs of the default workload (layer 0: qkv, out-proj, ffn1, ffn2 after 6 conv layers + projection)
timeout 900 $NCU --set full -k regex:gemm_tc2_kernel --launch-skip 7 -c 4 -f -o gpurun_out/r2c_ncu_dominant python bench.py --profile-step --steps 1 --warmup 0 > gpurun_out/r2c_ncu_dominant.log 2>&1
exp r2c_ncu_dominant
T="python bench.py --config TINY --clips 64 --seconds 10 --profile-step --steps 1 --warmup 0"
cap() { name=$1; regex=$2; skip=$3; cnt=$4; src=$5; timeout 600 $NCU --set full ${src:+--import-source on} -k regex:$regex --launch-skip $skip -c $cnt -f -o gpurun_out/r2c_ncu_$name $T > gpurun_out/r2c_ncu_$name.log 2>&1; exp r2c_ncu_$name $src; }
cap attn attn_tc_kernel 0 12 src
cap flame flame_tc_kernel 0 1 src
cap conv0 conv0 0 1
cap ln ln_affine_kernel 4 3
cap adaln adaln_kernel 0 10
cap bits "bits_|bsq_|argmax_bits" 0 12
cap pool "audio_pool|act_cast|savgol|motion_norm|audio_stats" 0 6
cap gemm1 "gemm_tc_kernel" 0 16
cap skinny skinny_gemm_kernel 0 4
timeout 600 $NCU --set full -k regex:"resample_mix|split_bf16|vertex_normals|ema_scan" -c 8 -f -o gpurun_out/r2c_ncu_misc python tools_kernels_once.py > gpurun_out/r2c_ncu_misc.log 2>&1
exp r2c_ncu_misc
rm -f gpurun_out/*.ncu-rep
# launch list of the bench command (one eager step of the default workload: the same kernels a graph replay runs)
timeout 1200 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2c_launches.csv python bench.py --profile-step --steps 1 --warmup 0 > gpurun_out/r2c_launches.log 2>&1
gzip -f gpurun_out/r2c_launches.csv
for f in gpurun_out/*.log; do tail -c 2000 $f > $f.tail; mv $f.tail $f; done
du -sh gpurun_out; ls -la gpurun_out | head -60
echo done
