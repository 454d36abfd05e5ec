"""Drop-in seam.  The reference has no FFI: its plug-in mechanism is YAML `target:` dotted paths
resolved by importlib (ldm/util.py:71-84; configs/test_*.yaml:9,27,43,64,76) plus a handful of
direct imports in inference.py.  `install()` registers this package's mirror modules in
`sys.modules` under the reference's own names, so an unmodified config / inference.py picks up the
B200 implementations of exactly the hot-path classes:

    ldm.util
    ldm.modules.attention
    ldm.modules.diffusionmodules.{openaimodel, text_grounding_net, util, convnext}
    ldm.models.diffusion.{plms, plms_instance, ldm, ddpm}
    grounding_input.text_grounding_tokinzer_input
    utils.model  (set_alpha_scale / alpha_generator only)

Anything else of the reference (autoencoder, CLIP encoders, dataset code) is not shadowed unless
`strict=False` and it is importable from the reference checkout on sys.path.
"""
from __future__ import annotations

import importlib
import sys
import types

_MAP = {
    "ldm.util": "instancediffusion_b200.ldm.util",
    "ldm.modules.attention": "instancediffusion_b200.ldm.modules.attention",
    "ldm.modules.diffusionmodules.openaimodel": "instancediffusion_b200.ldm.modules.diffusionmodules.openaimodel",
    "ldm.modules.diffusionmodules.text_grounding_net": "instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net",
    "ldm.modules.diffusionmodules.util": "instancediffusion_b200.ldm.modules.diffusionmodules.util",
    "ldm.modules.diffusionmodules.convnext": "instancediffusion_b200.ldm.modules.diffusionmodules.convnext",
    "ldm.models.diffusion.plms": "instancediffusion_b200.ldm.models.diffusion.plms",
    "ldm.models.diffusion.plms_instance": "instancediffusion_b200.ldm.models.diffusion.plms_instance",
    "ldm.models.diffusion.ldm": "instancediffusion_b200.ldm.models.diffusion.ldm",
    "ldm.models.diffusion.ddpm": "instancediffusion_b200.ldm.models.diffusion.ddpm",
    "grounding_input.text_grounding_tokinzer_input": "instancediffusion_b200.grounding_input.text_grounding_tokinzer_input",
}
_PKGS = {
    "ldm": "instancediffusion_b200.ldm",
    "ldm.modules": "instancediffusion_b200.ldm.modules",
    "ldm.modules.diffusionmodules": "instancediffusion_b200.ldm.modules.diffusionmodules",
    "ldm.models": "instancediffusion_b200.ldm.models",
    "ldm.models.diffusion": "instancediffusion_b200.ldm.models.diffusion",
    "grounding_input": "instancediffusion_b200.grounding_input",
}


def install(shadow_utils_model: bool = True) -> None:
    """Alias the mirror modules under the reference's import paths."""
    for alias, real in {**_PKGS, **_MAP}.items():
        sys.modules[alias] = importlib.import_module(real)
    if shadow_utils_model:
        from .utils import model as um
        pkg = sys.modules.get("utils")
        if pkg is None:
            pkg = types.ModuleType("utils")
            pkg.__path__ = []  # namespace-like
            sys.modules["utils"] = pkg
        shim = types.ModuleType("utils.model")
        shim.set_alpha_scale = um.set_alpha_scale
        shim.alpha_generator = um.alpha_generator
        sys.modules["utils.model"] = shim
        pkg.model = shim


def uninstall() -> None:
    for alias in list({**_PKGS, **_MAP}) + ["utils.model"]:
        sys.modules.pop(alias, None)
