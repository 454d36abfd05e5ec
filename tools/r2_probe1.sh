#!/bin/bash
# round-2 probe 1: which epilogue-warp configurations run at all (each launch in its own process + timeout),
# then tests / per-shape timings of the ones that do, then the attention exp2 variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_probe1.log
: > $L
for ew in 13 17 12 16; do
  echo "=== EW $ew" >> $L
  IDIFF_GEMM_EW=$ew timeout 200 python tools/run_one_gemm.py proj320,geglu320 2 >> $L 2>&1
done
ok16=$(awk '/=== EW 16/{f=1} f&&/rc 0/{c++} END{print c+0}' $L)
echo "ok16=$ok16" >> $L
if [ "$ok16" = "2" ]; then
  echo "=== tests EW16" >> $L
  timeout 400 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv or geglu" 2>&1 | tail -15 >> $L
  echo "=== bench EW16" >> $L
  timeout 300 python tools/bench_kernels.py ew16 gemm 2>&1 | tail -40 >> $L
fi
for poly in 0 2 3 4; do
  echo "=== att2 poly $poly" >> $L
  IDIFF_ATT2_POLY=$poly timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3 >> $L
  IDIFF_ATT2_POLY=$poly timeout 200 python tools/bench_kernels.py poly$poly attn 2>&1 | grep -E "N4096|N1024" >> $L
done
tail -30 $L
