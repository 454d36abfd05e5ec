"""Sweep the GEMM plan (tile width BN x {data-parallel, stream-K}, IDIFF_GEMM_PLAN) over every linear / conv shape
of the UNet forward (batch 8) and print the default plan's time next to the best forced one.
Usage: python tools/plan_sweep.py [batch]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instancediffusion_b200 import ops  # noqa: E402
from instancediffusion_b200.packing import pack_geglu  # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.sum()
        torch.cuda._sleep(300000)
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).half()


L = [(4096, 320), (1024, 640), (256, 1280)]
shapes = []
for (hw, C), nblk in zip(L, (5, 5, 5)):
    M = B * hw
    shapes += [(f"qkv C{C}", 2 * nblk, M, 3 * C, C, "lin"), (f"proj C{C}", 5 * nblk, M, C, C, "lin_res"),
               (f"geglu C{C}", 2 * nblk, M, 8 * C, C, "geglu"), (f"ff2 C{C}", 2 * nblk, M, C, 4 * C, "lin_res")]
shapes += [("qkv C1280@8", 2, B * 64, 3840, 1280, "lin"), ("geglu C1280@8", 2, B * 64, 10240, 1280, "geglu"),
           ("ff2 C1280@8", 2, B * 64, 1280, 5120, "lin_res"), ("proj C1280@8", 5, B * 64, 1280, 1280, "lin_res")]
convs = [("conv 320->320 @64", 4, 64, 320, 320), ("conv 640->320 @64", 2, 64, 640, 320), ("conv 960->320 @64", 1, 64, 960, 320),
         ("conv 640->640 @64", 1, 64, 640, 640), ("conv 640->640 @32", 6, 32, 640, 640), ("conv 320->640 @32", 1, 32, 320, 640),
         ("conv 1280->640 @32", 1, 32, 1280, 640), ("conv 1920->640 @32", 1, 32, 1920, 640), ("conv 960->640 @32", 1, 32, 960, 640),
         ("conv 1280->1280 @32", 1, 32, 1280, 1280), ("conv 1280->1280 @16", 9, 16, 1280, 1280), ("conv 640->1280 @16", 1, 16, 640, 1280),
         ("conv 2560->1280 @16", 2, 16, 2560, 1280), ("conv 1920->1280 @16", 1, 16, 1920, 1280),
         ("conv 1280->1280 @8", 11, 8, 1280, 1280), ("conv 2560->1280 @8", 3, 8, 2560, 1280)]

plans = [(bn, sk) for bn in (128, 160, 192, 256) for sk in (0, 1)]
tot_def = tot_best = 0.0
print(f"{'shape':24s} {'cnt':>3s} {'default':>8s} {'best':>8s}  best plan   all (bn,sk: us)")


def sweep(label, cnt, fn, geglu=False):
    global tot_def, tot_best
    os.environ.pop("IDIFF_GEMM_PLAN", None)
    t_def = timed(fn)
    res = {}
    for bn, sk in plans:
        if geglu and bn != 256:
            continue
        os.environ["IDIFF_GEMM_PLAN"] = f"{bn},{sk}"
        try:
            res[(bn, sk)] = timed(fn)
        except Exception as e:  # a plan the kernel refuses
            res[(bn, sk)] = float("inf")
    os.environ.pop("IDIFF_GEMM_PLAN", None)
    best = min(res, key=res.get)
    tot_def += cnt * t_def
    tot_best += cnt * min(t_def, res[best])
    print(f"{label:24s} {cnt:3d} {t_def:8.1f} {res[best]:8.1f}  {best}   " +
          " ".join(f"{k[0]},{k[1]}:{v:.0f}" for k, v in res.items()), flush=True)


for label, cnt, M, N, K, kind in shapes:
    a = rnd(M, K)
    w = rnd(N, K, scale=1 / math.sqrt(K))
    bias = torch.randn(N, device=dev)
    if kind == "geglu":
        wp, bp = pack_geglu(w, bias)
        out = torch.empty((M, N // 2), dtype=torch.float16, device=dev)
        sweep(label, cnt, lambda: ops.gemm(a, wp, bp, geglu=True, out=out), geglu=True)
    elif kind == "lin_res":
        res_ = rnd(M, N)
        out = torch.empty((M, N), dtype=torch.float16, device=dev)
        sweep(label, cnt, lambda: ops.gemm(a, w, bias, residual=res_, out=out))
    else:
        out = torch.empty((M, N), dtype=torch.float16, device=dev)
        sweep(label, cnt, lambda: ops.gemm(a, w, bias, out=out))
for label, cnt, hw, cin, cout in convs:
    M = B * hw * hw
    a = rnd(M, cin)
    w = rnd(cout, 9 * cin, scale=1 / math.sqrt(9 * cin))
    bias = torch.randn(cout, device=dev)
    res_ = rnd(M, cout)
    out = torch.empty((M, cout), dtype=torch.float16, device=dev)
    sweep(label, cnt, lambda: ops.gemm(a, w, bias, conv=(B, hw, hw, cin), residual=res_, out=out))
print(f"count-weighted per forward: default {tot_def / 1e3:.2f} ms, best-of-sweep {tot_best / 1e3:.2f} ms")
