"""Launch a few representative GEMM shapes (for ncu --set full captures).
Usage: python tools/run_one_gemm.py qkv320|proj320|geglu320|conv320|conv1280 [n]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instancediffusion_b200 import ops
from instancediffusion_b200.packing import pack_geglu

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if "," in which:  # several shapes in one process, progress printed after each (hang localisation)
    import subprocess
    for w_ in which.split(","):
        print("shape", w_, flush=True)
        try:
            rc = subprocess.run([sys.executable, __file__, w_, str(n)], timeout=60).returncode
        except subprocess.TimeoutExpired:
            rc = "TIMEOUT (hang)"
        print("  rc", rc, flush=True)
    sys.exit(0)
dev = torch.device("cuda:0")
B = 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
if which == "qkv320":
    a, w = r(B * 4096, 320), r(960, 320, sc=0.05)
    fn = lambda: ops.gemm(a, w)
elif which == "proj320":
    a, w, res, b = r(B * 4096, 320), r(320, 320, sc=0.05), r(B * 4096, 320), torch.randn(320, device=dev)
    fn = lambda: ops.gemm(a, w, b, residual=res)
elif which == "geglu320":
    a, w, b = r(B * 4096, 320), r(2560, 320, sc=0.05), torch.randn(2560, device=dev)
    wp, bp = pack_geglu(w, b)
    fn = lambda: ops.gemm(a, wp, bp, geglu=True)
elif which == "proj1280":
    a, w, res, b = r(B * 256, 1280), r(1280, 1280, sc=0.03), r(B * 256, 1280), torch.randn(1280, device=dev)
    fn = lambda: ops.gemm(a, w, b, residual=res)
elif which == "conv320":
    a, w, b, res = r(B * 4096, 320), r(320, 2880, sc=0.02), torch.randn(320, device=dev), r(B * 4096, 320)
    fn = lambda: ops.gemm(a, w, b, conv=(B, 64, 64, 320), residual=res)
elif which == "conv1280":
    a, w, b, res = r(B * 256, 1280), r(1280, 11520, sc=0.01), torch.randn(1280, device=dev), r(B * 256, 1280)
    fn = lambda: ops.gemm(a, w, b, conv=(B, 16, 16, 1280), residual=res)
for _ in range(n):
    fn()
torch.cuda.synchronize()
print("ok")
