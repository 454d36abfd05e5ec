"""ConvNeXt-tiny parameter container (ldm/modules/diffusionmodules/convnext.py:15-123): depths
3-3-9-3, dims 96/192/384/768 -- the mask encoder inside UniFusion.  It holds the 178 state_dict
keys so reference checkpoints load strict.  It only runs when a sample carries non-zero `segs`
(mask conditioning), at most once per sample after hoisting; SURVEY.md section 8(f) lists it as a
"next" row, so `forward` fails loudly instead of silently falling back to library kernels."""
import torch
import torch.nn as nn


class LayerNorm(nn.Module):
    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps
        self.data_format = data_format
        self.normalized_shape = (normalized_shape,)


class Block(nn.Module):
    def __init__(self, dim, drop_path=0., layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones((dim)), requires_grad=True)


class ConvNeXt(nn.Module):
    def __init__(self, in_chans=3, num_classes=1000, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768),
                 drop_path_rate=0., layer_scale_init_value=1e-6, head_init_scale=1.):
        super().__init__()
        self.downsample_layers = nn.ModuleList()
        self.downsample_layers.append(nn.Sequential(
            nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4),
            LayerNorm(dims[0], eps=1e-6, data_format="channels_first")))
        for i in range(3):
            self.downsample_layers.append(nn.Sequential(
                LayerNorm(dims[i], eps=1e-6, data_format="channels_first"),
                nn.Conv2d(dims[i], dims[i + 1], kernel_size=2, stride=2)))
        self.stages = nn.ModuleList([
            nn.Sequential(*[Block(dim=dims[i], layer_scale_init_value=layer_scale_init_value) for _ in range(depths[i])])
            for i in range(4)])

    def forward(self, x):
        raise NotImplementedError(
            "ConvNeXt mask encoder (non-zero `segs`) is not yet implemented in instancediffusion_b200 "
            "(SURVEY.md section 8f 'next'); box/point/scribble conditioning does not need it")


def convnext_tiny(pretrained=False, in_22k=False, **kwargs):
    """No network download here (the reference fetches ImageNet weights at construction,
    convnext.py:152-158); weights arrive with the InstanceDiffusion checkpoint."""
    return ConvNeXt(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], **kwargs)
