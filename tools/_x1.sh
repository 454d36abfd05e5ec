export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv" 2>&1 | tail -2
timeout 300 python tools/plan_sweep.py 8 > gpurun_out/r2_plan_sweep_b8_new.log 2>&1; tail -1 gpurun_out/r2_plan_sweep_b8_new.log
timeout 300 python tools/plan_sweep.py 16 > gpurun_out/r2_plan_sweep_b16_new.log 2>&1; awk '{print $1,$2,$3,$4,$5,$6,$7,$8}' gpurun_out/r2_plan_sweep_b16_new.log | column -t | cut -c1-100
