export PYTHONUNBUFFERED=1
timeout 700 python -m pytest tests/test_parity_r2_gpu.py tests/test_modules_gpu.py tests/test_masked_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_x3.json 2> gpurun_out/r2_bench_x3.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_x3.json'))
print(d['value'], d['e2e']['value'], d['clocks'], d['mis036']['value'], d['mis036']['clocks'])
P
tail -3 gpurun_out/r2_bench_x3.err
