export PYTHONUNBUFFERED=1
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_final4.json 2> gpurun_out/r2_bench_final4.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_final4.json'))
print(d['value'], d['e2e']['value'], d['clocks'], '| mis', d['mis036']['value'], d['mis036']['e2e'], d['mis036']['clocks']['sm_mhz'], d['roofline']['frac'], d['model_roofline']['frac_of_sustained_peak'])
P
tail -2 gpurun_out/r2_bench_final4.err
