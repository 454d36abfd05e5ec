"""One eager batched (cond+uncond, batch 8) UNet forward between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Usage: python tools/profile_forward.py [batch] [alpha]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from instancediffusion_b200 import synthetic
from instancediffusion_b200.utils.model import set_alpha_scale
from instancediffusion_b200.weights import build_unet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
alpha = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda:0")
model = build_unet("box", dev, seed=0)
model.use_cuda_graph = False
gti = model.grounding_tokenizer_input
inp, uc = synthetic.make_sampler_inputs(gti, B, 8, 77, "box", mis=False, device=dev)
inp["timesteps"] = torch.full((B,), 601, dtype=torch.long, device=dev)
un = dict(x=inp["x"], timesteps=inp["timesteps"], context=uc)
set_alpha_scale(model, alpha)
with torch.no_grad():
    model.forward_batched([inp, un])
    torch.cuda.synchronize()
    if len(sys.argv) > 3 and sys.argv[3] == "cold":  # include the per-sample hoisted work (UniFusion tokens, object K/V)
        model.clear_caches()
        model.use_cuda_graph = False
    torch.cuda.profiler.start()
    model.forward_batched([inp, un])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
