"""Which operand's L2 -> SM traffic bounds the conv / GEMM mainloop?  Times a few shapes with A and / or B
tile loads switched off (IDIFF_GEMM_SKIP; results are garbage, only the timing matters)."""
import math, os, subprocess, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from instancediffusion_b200 import ops
    dev = torch.device("cuda:0")
    B = 8
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    def timed(fn, iters=5):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(iters):
            flush.sum(); torch.cuda._sleep(300000)
            s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e) * 1e3)
        ts.sort(); return ts[len(ts) // 2]
    r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
    out = []
    for label, hw, cin, cout in [("conv320@64", 64, 320, 320), ("conv640@32", 32, 640, 640), ("conv1280@16", 16, 1280, 1280), ("conv1280@32", 32, 1280, 1280)]:
        M = B * hw * hw
        a, w, b, res = r(M, cin), r(cout, 9 * cin, sc=0.02), torch.randn(cout, device=dev), r(M, cout)
        o = torch.empty((M, cout), dtype=torch.float16, device=dev)
        out.append((label, timed(lambda: ops.gemm(a, w, b, conv=(B, hw, hw, cin), residual=res, out=o))))
    for label, M, N, K in [("geglu320", 32768, 2560, 320), ("qkv640", 8192, 1920, 640), ("ff2_640", 8192, 640, 2560)]:
        a, w = r(M, K), r(N, K, sc=0.03)
        o = torch.empty((M, N), dtype=torch.float16, device=dev)
        out.append((label, timed(lambda: ops.gemm(a, w, out=o))))
    print(" ".join(f"{l}={t:.1f}" for l, t in out))
else:
    for skip in (0, 1, 2, 3):
        env = dict(os.environ, IDIFF_GEMM_SKIP=str(skip))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True, timeout=200)
        print(f"skip={skip} (bit0 no A, bit1 no B): {r.stdout.strip()} {r.stderr.strip()[-300:]}", flush=True)
