"""Uninitialised-read probe: poison the caching allocator's free blocks with NaN, run the MIS sampler eagerly with
every ops.* wrapper checking its outputs, report the first op whose output is non-finite while its inputs were finite."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import test_parity_r2_gpu as P
from instancediffusion_b200 import ops
from instancediffusion_b200.weights import build_unet
dev = torch.device("cuda:0")
unet = build_unet("box", dev, seed=0)
unet._sd_conv = torch.load(os.path.join(ROOT, "tests/golden/sd15_first_conv.pt"), map_location="cpu")
gold = torch.load(os.path.join(ROOT, "tests/golden/samplers_extra.pt"), map_location="cpu")
name = "mis_S10_n3"
sc = cases.SAMPLER_EXTRA_CASES[name]

def poison():
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    big = [torch.full((1 << 28,), float("nan"), device=dev) for _ in range(24)]  # 24 GB of NaN
    torch.cuda.synchronize()
    del big

found = [0]
def tensors(o):
    if isinstance(o, torch.Tensor): return [o]
    if isinstance(o, (list, tuple)): return [t for x in o for t in tensors(x)]
    if hasattr(o, "t") and isinstance(getattr(o, "t"), torch.Tensor): return [o.t]
    return []
def wrap(fname):
    fn = getattr(ops, fname)
    def w(*a, **k):
        ins = tensors(list(a)) + tensors(list(k.values()))
        out = fn(*a, **k)
        if found[0] < 6:
            bad_out = [t for t in tensors(out) if t.is_floating_point() and not torch.isfinite(t.float()).all()]
            if bad_out:
                bad_in = [t for t in ins if t.is_floating_point() and not torch.isfinite(t.float()).all()]
                tag = "PROPAGATED" if bad_in else "ORIGIN"
                shapes = [tuple(t.shape) for t in ins]
                print(f"[{tag}] ops.{fname}: non-finite output {[tuple(t.shape) for t in bad_out]} inputs {shapes} kwargs {[kk for kk in k]}", flush=True)
                if not bad_in: found[0] += 1
        return out
    setattr(ops, fname, w)
for f in ["gemm", "attention", "groupnorm", "layernorm", "scaleu_concat", "nchw_f32_to_nhwc_f16", "upsample_nearest2x",
          "im2col_s2", "fourier_embed", "timestep_embedding", "silu", "row_stats", "latent_mean"]:
    wrap(f)

os.environ["IDIFF_CUDA_GRAPH"] = "0"
orig_run = P._run_sampler
def run(poisoned):
    if poisoned: poison()
    # _run_sampler forces use_cuda_graph=True: patch the attribute after it sets it via a property-free trick
    unet.__class__.use_cuda_graph = property(lambda self: False, lambda self, v: None)
    x = orig_run(unet, sc, dev).float().cpu()
    print("poisoned" if poisoned else "clean", "finite:", bool(torch.isfinite(x).all()),
          "rel to golden %.4e" % ((x - gold[name]).norm() / gold[name].norm()).item(), flush=True)
run(False)
run(True)
