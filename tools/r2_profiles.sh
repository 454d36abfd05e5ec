#!/bin/bash
# Round-2 profile captures (run under gpurun): launch list of one eager forward + one `--set full` capture per
# kernel.  Summaries are extracted afterwards with tools/ncu_summary.py / tools/ncu_traffic.py.
# Usage: bash tools/r2_profiles.sh [all|quick]   (quick: launch list + the kernels changed last)
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
mode=${1:-all}
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --profile-from-start off --csv --log-file gpurun_out/r2_launches_raw.csv python tools/profile_forward.py 4 1.0 > gpurun_out/r2_prof.log 2>&1
if [ "$mode" = all ]; then
  kernels="gn_fused_kernel gn_stats_kernel gn_apply_kernel scaleu_coef_kernel scaleu_apply_kernel layernorm40_kernel fourier_embed_kernel attention_kernel<\(int\)80 attention_kernel<\(int\)160 attention2_kernel"
else
  kernels="gn_fused_kernel attention_kernel<\(int\)80"
fi
for k in $kernels; do
  tag=$(echo $k | tr -cd 'a-z0-9_')
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k "regex:$k" -c 1 -f \
    -o gpurun_out/r2k_$tag python tools/profile_forward.py 4 1.0 cold >> gpurun_out/r2_prof.log 2>&1
done
for g in geglu320 proj320 conv320; do
  [ "$mode" = quick ] && [ $g != geglu320 ] && continue
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 2 -c 1 -f -o gpurun_out/r2f_$g python tools/run_one_gemm.py $g 3 >> gpurun_out/r2_prof.log 2>&1
done
ls -la gpurun_out/r2k_* gpurun_out/r2f_* gpurun_out/r2_launches_raw.csv
tail -3 gpurun_out/r2_prof.log
