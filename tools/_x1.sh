export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_parity_r2_gpu.py tests/test_bf16_gpu.py -m gpu -q -s -x > gpurun_out/r2_parity_fix.log 2>&1; tail -3 gpurun_out/r2_parity_fix.log; grep -h "envelope\]\|sampler_extra/" gpurun_out/r2_parity_fix.log
