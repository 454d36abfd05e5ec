"""TEST INFRASTRUCTURE.  Import the *unmodified* reference modules (frank-xwang/InstanceDiffusion)
on the CPU so they can serve as the parity oracle and golden-vector generator (SURVEY.md
appendix C).  Nothing from the reference is copied: the modules are imported from where they lie.
"""
from __future__ import annotations

import os
import sys
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def reference_root() -> str | None:
    for cand in (os.environ.get("IDIFF_REF"), "/root/reference", os.path.join(REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "ldm")):
            return cand
    return None


_imported = None


def import_reference():
    """Returns a namespace with the reference classes.  Side effects (this process only): the
    reference root is put first on sys.path and becomes the cwd (openaimodel.py:476 opens
    pretrained/... relative to it); `timm` and torch.hub are stubbed (convnext.py:12-13,156)."""
    global _imported
    if _imported is not None:
        return _imported
    root = reference_root()
    if root is None:
        raise RuntimeError("reference checkout not found (set IDIFF_REF)")
    if any(k == "ldm" or k.startswith("ldm.") for k in sys.modules):
        raise RuntimeError("an `ldm` package is already imported (dropin.install() active?); "
                           "the oracle must run in a process that has not shadowed the reference")
    sys.path.insert(0, root)
    os.chdir(root)
    stubs = {n: types.ModuleType(n) for n in ("timm", "timm.models", "timm.models.layers", "timm.models.registry")}
    stubs["timm.models.layers"].trunc_normal_ = torch.nn.init.trunc_normal_
    stubs["timm.models.layers"].DropPath = type(
        "DropPath", (torch.nn.Identity,), {"__init__": lambda s, p=0.: torch.nn.Identity.__init__(s)})
    stubs["timm.models.registry"].register_model = lambda f: f
    sys.modules.update(stubs)
    torch.hub.load_state_dict_from_url = lambda *a, **k: {"model": {}}

    import warnings
    warnings.filterwarnings("ignore")
    ns = types.SimpleNamespace()
    from ldm.modules import attention as att
    from ldm.modules.diffusionmodules import openaimodel as oai
    from ldm.modules.diffusionmodules import text_grounding_net as tgn
    from ldm.modules.diffusionmodules import util as dutil
    from ldm.models.diffusion.ldm import LatentDiffusion
    from ldm.models.diffusion.plms import PLMSSampler
    from ldm.models.diffusion.plms_instance import PLMSSamplerInst
    from grounding_input.text_grounding_tokinzer_input import GroundingNetInput
    ns.root = root
    ns.attention, ns.openaimodel, ns.text_grounding_net, ns.util = att, oai, tgn, dutil
    ns.LatentDiffusion, ns.PLMSSampler, ns.PLMSSamplerInst, ns.GroundingNetInput = (
        LatentDiffusion, PLMSSampler, PLMSSamplerInst, GroundingNetInput)
    _imported = ns
    return ns


class fast_init:
    """Skip the (slow, overwritten anyway) default parameter initialisation while constructing
    reference modules: the reference cannot be built on the meta device (convnext.py:86 calls
    .item() at construction)."""
    _names = ("kaiming_uniform_", "uniform_", "trunc_normal_", "normal_", "constant_", "zeros_", "ones_")

    def __enter__(self):
        self._saved = {n: getattr(torch.nn.init, n) for n in self._names}
        for n in self._names:
            setattr(torch.nn.init, n, lambda t, *a, **k: t)
        return self

    def __exit__(self, *exc):
        for n, f in self._saved.items():
            setattr(torch.nn.init, n, f)
        return False


def set_alpha_scale(ref, model, alpha_scale):
    """utils/model.py:78-81 restated (utils/model.py itself needs omegaconf at import time)."""
    for module in model.modules():
        if type(module) == ref.attention.GatedSelfAttentionDense:
            module.scale = alpha_scale


def build_ref_unet(ref, flavor: str = "box", seed: int = 0):
    """The reference UNetModel built from the configs/test_*.yaml parameters, filled with the same
    synthetic weights the CUDA path loads (instancediffusion_b200/weights.py)."""
    from instancediffusion_b200.weights import synth_tensor, unet_config
    cfg = unet_config(flavor, tokenizer_target="ldm.modules.diffusionmodules.text_grounding_net.UniFusion")
    with fast_init():
        model = ref.openaimodel.UNetModel(**cfg).eval()
    sd = {k: synth_tensor(k, tuple(v.shape), seed) for k, v in model.state_dict().items()}
    model.load_state_dict(sd, strict=True)
    model.grounding_tokenizer_input = ref.GroundingNetInput()
    return model
