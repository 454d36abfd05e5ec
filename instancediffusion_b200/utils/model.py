"""Host helpers the samplers are constructed with (utils/model.py:78-117 of the reference)."""
import numpy as np

from ..ldm.modules.attention import GatedSelfAttentionDense


def set_alpha_scale(model, alpha_scale):
    """utils/model.py:78-81: write `.scale` on every gated fuser (exact type match, as there)."""
    for module in model.modules():
        if type(module) == GatedSelfAttentionDense:
            module.scale = alpha_scale


def alpha_generator(length, type=None):
    """utils/model.py:83-117: [alpha=1 stage | linear decay stage | alpha=0 stage] fractions."""
    if type is None:
        type = [1, 0, 0]
    assert len(type) == 3
    assert type[0] + type[1] + type[2] == 1
    n0 = int(type[0] * length)
    n1 = int(type[1] * length)
    n2 = length - n0 - n1
    decay = list(np.arange(start=0, stop=1, step=1 / n1)[::-1]) if n1 != 0 else []
    alphas = [1] * n0 + decay + [0] * n2
    assert len(alphas) == length
    return alphas
