"""Bisect the first-run / later-run difference of the MIS sampler."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import test_parity_r2_gpu as P
from instancediffusion_b200 import ops
from instancediffusion_b200.weights import build_unet
dev = torch.device("cuda:0")
gold = torch.load(os.path.join(ROOT, "tests/golden/samplers_extra.pt"), map_location="cpu")
name = "mis_S10_n3"
sc = cases.SAMPLER_EXTRA_CASES[name]
sdconv = torch.load(os.path.join(ROOT, "tests/golden/sd15_first_conv.pt"), map_location="cpu")

trace = []
_pu = ops.plms_update
def pu(x, e_c, e_u, *a, **k):
    trace.append((float(x.double().sum()), float(e_c.double().sum()), float(e_u.double().sum()) if e_u is not None else 0.0))
    return _pu(x, e_c, e_u, *a, **k)
ops.plms_update = pu
import instancediffusion_b200.ldm.models.diffusion._plms_common as PC
assert PC.ops is ops

def new_model():
    m = build_unet("box", dev, seed=0)
    m._sd_conv = sdconv
    return m

def run(m, tag):
    trace.clear()
    x = P._run_sampler(m, sc, dev).float().cpu()
    print(tag, "rel to golden %.4e" % ((x - gold[name]).norm() / gold[name].norm()).item(), "calls", len(trace), flush=True)
    return list(trace)

m1 = new_model()
t0 = run(m1, "model1 run0")
t1 = run(m1, "model1 run1")
for i, (a, b) in enumerate(zip(t0, t1)):
    if a != b:
        print("first differing plms_update call:", i, "of", len(t0), " x equal:", a[0] == b[0], " e_c equal:", a[1] == b[1], " e_u equal:", a[2] == b[2])
        break
m1.invalidate_pack()
run(m1, "model1 run2 after invalidate_pack")
m2 = new_model()
run(m2, "model2 (fresh object, same process) run0")
run(m2, "model2 run1")
