export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_modules_gpu.py -m gpu -q -x -k "geglu or feedforward or FeedForward or ff or transformer or groupnorm" 2>&1 | tail -3
timeout 200 python tools/bench_kernels.py geglu2 gemm > gpurun_out/r2_geglu2.log 2>&1; grep -i "geglu" gpurun_out/r2_geglu2.log; tail -1 gpurun_out/r2_geglu2.log
timeout 100 python tools/bench_kernels.py gnf0 norm > gpurun_out/r2_gnf0.log 2>&1
IDIFF_GN_FUSED=1 timeout 100 python tools/bench_kernels.py gnf1 norm > gpurun_out/r2_gnf1.log 2>&1
paste <(grep groupnorm gpurun_out/r2_gnf0.log) <(grep groupnorm gpurun_out/r2_gnf1.log | cut -c30-)
