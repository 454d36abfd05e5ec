"""Determinism probe: the same sampler case several times in one process, with and without other work in between."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import test_parity_r2_gpu as P
from instancediffusion_b200.weights import build_unet
dev = torch.device("cuda:0")
unet = build_unet("box", dev, seed=0)
unet._sd_conv = torch.load(os.path.join(ROOT, "tests/golden/sd15_first_conv.pt"), map_location="cpu")
gold = torch.load(os.path.join(ROOT, "tests/golden/samplers_extra.pt"), map_location="cpu")
name = sys.argv[1] if len(sys.argv) > 1 else "mis_S10_n3"
sc = cases.SAMPLER_EXTRA_CASES[name]
outs = []
for i in range(4):
    if i == 2:  # other work in between: an eager forward of a different shape
        spec = cases.UNET_EXTRA_CASES["b4n8"]
        inp, uc, ts = P._inputs(unet, spec, dev)
        unet.clear_caches(); unet.use_cuda_graph = False
        unet(dict(x=inp["x"], timesteps=ts, context=inp["context"], grounding_input=inp["grounding_input"]))
    if i == 3:
        os.environ["IDIFF_CUDA_GRAPH_OFF"] = "1"
    x = P._run_sampler(unet, sc, dev).float().cpu()
    outs.append(x)
    print(i, "rel to golden %.4e" % ((x - gold[name]).norm() / gold[name].norm()).item(),
          "equal to run 0:", torch.equal(x, outs[0]), "rel to run 0 %.3e" % ((x - outs[0]).norm() / outs[0].norm()).item(), flush=True)
