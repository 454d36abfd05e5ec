export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention or groupnorm" 2>&1 | tail -3
timeout 100 python tools/bench_kernels.py att64 attn > gpurun_out/r2_att64.log 2>&1
IDIFF_ATT_BKV=128 timeout 100 python tools/bench_kernels.py att128 attn > gpurun_out/r2_att128.log 2>&1
paste <(grep -E "N1024|N256|N64" gpurun_out/r2_att64.log) <(grep -E "N1024|N256|N64" gpurun_out/r2_att128.log | cut -c30-)
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mis-leg > gpurun_out/r2_bench_x2.json 2> gpurun_out/r2_bench_x2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_x2.json'))
print(d['value'], d['e2e']['value'], d['clocks'])
for k,v in d['breakdown'].items(): print(k, round(v['ms'],3), v['launches'])
P
tail -3 gpurun_out/r2_bench_x2.err
