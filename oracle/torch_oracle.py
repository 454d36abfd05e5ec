"""TEST INFRASTRUCTURE -- a plain-PyTorch fp32 restatement of the reference's sampling hot path,
written functionally over a `state_dict` (no nn.Module, nothing imported from the reference, so it
runs on the GPU box where /root/reference does not exist).  Each function cites the reference
file:line it follows.  It is pinned against fixtures produced by the reference's own modules
(oracle/make_golden.py -> tests/golden/*.pt, checked by tests/test_oracle_cpu.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` leg may
import this file.  The product never does: instancediffusion_b200 has no CPU path.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
HEADS = 8  # configs/test_*.yaml:19 num_heads


# ------------------------------------------------------------------------------------------------
# ldm/modules/attention.py
# ------------------------------------------------------------------------------------------------
def linear(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def feed_forward(sd: SD, p: str, x):
    """FeedForward(glu=True), attention.py:36-63: Linear(C,8C) -> a*gelu(g) (erf) -> Linear(4C,C)."""
    h = linear(sd, p + ".net.0.proj", x)
    a, g = h.chunk(2, dim=-1)
    return linear(sd, p + ".net.2", a * F.gelu(g))


# False: explicit softmax(q k^T) v (bit-stable across torch builds, used for the CPU goldens' pin).
# True: F.scaled_dot_product_attention -- a second, equally valid fp16 realisation of the reference for the
# measured fp16 envelope (tests/test_parity_r2_gpu.py).
FUSED_SDPA = False


def _sdpa(q, k, v, heads, att_mask=None):
    """attention.py:130-144 / 183-185,257-267: head split, softmax(q k^T d^-1/2) v, head merge.  att_mask
    (B,1,N,N): entries <= 0 are filled with -inf before the softmax (:276-277, the non-efficient path)."""
    B, N, C = q.shape
    M = k.shape[1]
    d = C // heads
    q = q.view(B, N, heads, d).permute(0, 2, 1, 3)
    k = k.view(B, M, heads, d).permute(0, 2, 1, 3)
    v = v.view(B, M, heads, d).permute(0, 2, 1, 3)
    if att_mask is not None:
        sim = (q @ k.transpose(-1, -2)) * (d ** -0.5)
        sim = sim.masked_fill(att_mask <= 0.0, float("-inf"))
        return (torch.softmax(sim, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B, N, C)
    if FUSED_SDPA:  # what attention.py:139-143,262-266 calls; on a GPU under autocast this is the flash kernel
        return F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B, N, C)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(B, N, C)


def self_attention(sd: SD, p: str, x, heads=HEADS, att_mask=None):
    """SelfAttention.forward, attention.py:174-282 (att_mask: the instance-isolation mask of :187-255)."""
    o = _sdpa(linear(sd, p + ".to_q", x), linear(sd, p + ".to_k", x), linear(sd, p + ".to_v", x), heads, att_mask)
    return linear(sd, p + ".to_out.0", o)


def instance_attention_mask(att_masks, N: int):
    """attention.py:203-251 restated.  att_masks (B, n_objs, s, s) binary -> (B, 1, N, N) with N = s*s + 4*n_objs + 64:
    two visual tokens see each other iff some instance covers both; box / mask object tokens see (and are seen by)
    the pixels of their instance, point / scribble tokens and the trailing 64 tokens everything; 1e-9 on the
    diagonal keeps every row alive."""
    B, n, s1, s2 = att_masks.shape
    wh = s1 * s2
    m = att_masks.reshape(B, n, wh).float()
    both = torch.einsum("bki,bkj->bij", m, m)                 # number of instances covering both pixels
    out = torch.ones((B, 1, N, N), dtype=att_masks.dtype)
    out[:, 0, :wh, :wh] = (both >= 1.0).to(att_masks.dtype)
    rep = m.repeat(1, 4, 1)                                    # [box | point | scribble | mask] token rows
    out[:, 0, wh:N - 64, :wh] = rep
    out[:, 0, wh + n:wh + 3 * n, :wh] = 1
    out[:, 0, :wh, wh:N - 64] = rep.transpose(1, 2)
    out[:, 0, :wh, wh + n:wh + 3 * n] = 1
    return out + torch.eye(N, dtype=att_masks.dtype).view(1, 1, N, N) * 1e-9


def cross_attention(sd: SD, p: str, x, ctx, heads=HEADS):
    """CrossAttention.forward, attention.py:120-157."""
    o = _sdpa(linear(sd, p + ".to_q", x), linear(sd, p + ".to_k", ctx), linear(sd, p + ".to_v", ctx), heads)
    return linear(sd, p + ".to_out.0", o)


def layer_norm(sd: SD, p: str, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def gated_self_attention(sd: SD, p: str, x, objs, scale=1.0, heads=HEADS, att_masks=None, drop_box_mask=False):
    """GatedSelfAttentionDense.forward, attention.py:304-311 -- as written there: concatenate, norm,
    attend over all N+184 rows, keep the first N.  att_masks (B, n_objs, 64, 64): the efficient_attention=False
    path with the instance-isolation mask (:190-201: only at 64x64, only if some mask is set and boxes are kept)."""
    n_vis = x.shape[1]
    o = linear(sd, p + ".linear", objs)
    att_mask = None
    if att_masks is not None:
        N = n_vis + o.shape[1]
        if N - att_masks.shape[1] * 4 - 64 == 64 * 64 and float(att_masks.sum()) > 0.0 and not drop_box_mask:
            att_mask = instance_attention_mask(att_masks, N)
    a = self_attention(sd, p + ".attn", layer_norm(sd, p + ".norm1", torch.cat([x, o], dim=1)), heads, att_mask)
    x = x + scale * torch.tanh(sd[p + ".alpha_attn"]) * a[:, :n_vis]
    x = x + scale * torch.tanh(sd[p + ".alpha_dense"]) * feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm2", x))
    return x


def basic_transformer_block(sd: SD, p: str, x, ctx, objs, scale=1.0, heads=HEADS):
    """BasicTransformerBlock._forward, attention.py:333-338."""
    x = self_attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), heads) + x
    x = gated_self_attention(sd, p + ".fuser", x, objs, scale, heads)
    x = cross_attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), ctx, heads) + x
    x = feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    return x


def spatial_transformer(sd: SD, p: str, x, ctx, objs, scale=1.0, heads=HEADS):
    """SpatialTransformer.forward, attention.py:366-379 (GroupNorm eps 1e-6, :75-76)."""
    b, c, h, w = x.shape
    x_in = x
    x = F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    x = F.conv2d(x, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    i = 0
    while f"{p}.transformer_blocks.{i}.norm1.weight" in sd:
        x = basic_transformer_block(sd, f"{p}.transformer_blocks.{i}", x, ctx, objs, scale, heads)
        i += 1
    x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + x_in


# ------------------------------------------------------------------------------------------------
# ldm/modules/diffusionmodules/openaimodel.py
# ------------------------------------------------------------------------------------------------
def resblock(sd: SD, p: str, x, emb):
    """ResBlock._forward, openaimodel.py:237-257 (no up/down, no scale-shift)."""
    h = F.silu(F.group_norm(x.float(), 32, sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"], 1e-5))
    h = F.conv2d(h, sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = linear(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + e[:, :, None, None]
    h = F.silu(F.group_norm(h, 32, sd[p + ".out_layers.0.weight"], sd[p + ".out_layers.0.bias"], 1e-5))
    h = F.conv2d(h, sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def upsample(sd: SD, p: str, x):
    """Upsample.forward, openaimodel.py:100-110."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    return F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)


def downsample(sd: SD, p: str, x):
    """Downsample.forward, openaimodel.py:130-141 (stride 2, padding 1)."""
    return F.conv2d(x, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)


def fourier_filter(x, threshold, scale):
    """Fourier_filter, openaimodel.py:25-48, as written there (FFT, centred mask, inverse, real part)."""
    B, C, H, W = x.shape
    xf = torch.fft.fftshift(torch.fft.fftn(x.float(), dim=(-2, -1)), dim=(-2, -1))
    mask = torch.ones((B, C, H, W), device=x.device)
    crow, ccol = H // 2, W // 2
    mask[..., crow - threshold:crow + threshold, ccol - threshold:ccol + threshold] = scale
    xf = torch.fft.ifftshift(xf * mask, dim=(-2, -1))
    return torch.fft.ifftn(xf, dim=(-2, -1)).real.to(x.dtype)


def timestep_embedding(t, dim, max_period=10000):
    """util.py:160-180: [cos(t f) | sin(t f)], f_k = exp(-ln(max_period) k / half)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ------------------------------------------------------------------------------------------------
# ldm/modules/diffusionmodules/text_grounding_net.py
# ------------------------------------------------------------------------------------------------
def fourier_embed(x, num_freqs=16, temperature=100):
    """FourierEmbedder.__call__, util.py:19-26: cat_k [sin(f_k x), cos(f_k x)], f_k = T^(k/num)."""
    freqs = temperature ** (torch.arange(num_freqs) / num_freqs)
    out = []
    for f in freqs:
        out.append(torch.sin(f * x))
        out.append(torch.cos(f * x))
    return torch.cat(out, dim=-1)


def convnext_block(sd: SD, p: str, x):
    """convnext.py:38-51: dwconv7x7 -> LayerNorm(C, eps 1e-6) over channels -> Linear(C,4C) -> GELU ->
    Linear(4C,C) -> * gamma -> + input."""
    C = x.shape[1]
    h = F.conv2d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=3, groups=C).permute(0, 2, 3, 1)
    h = F.layer_norm(h, (C,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    h = linear(sd, p + ".pwconv2", F.gelu(linear(sd, p + ".pwconv1", h)))
    if p + ".gamma" in sd:
        h = sd[p + ".gamma"] * h
    return x + h.permute(0, 3, 1, 2)


def _ln_channels_first(sd: SD, p: str, x, eps=1e-6):
    """convnext.py:139-144."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][:, None, None] * x + sd[p + ".bias"][:, None, None]


def convnext(sd: SD, p: str, x, depths=(3, 3, 9, 3)):
    """ConvNeXt.forward_features, convnext.py:107-111 (stem :71-74, downsamplers :77-81, stages :86-92)."""
    for i in range(4):
        d = f"{p}.downsample_layers.{i}"
        if i == 0:
            x = F.conv2d(x, sd[d + ".0.weight"], sd[d + ".0.bias"], stride=4)
            x = _ln_channels_first(sd, d + ".1", x)
        else:
            x = _ln_channels_first(sd, d + ".0", x)
            x = F.conv2d(x, sd[d + ".1.weight"], sd[d + ".1.bias"], stride=2)
        for j in range(depths[i]):
            x = convnext_block(sd, f"{p}.stages.{i}.{j}", x)
    return x


def _mlp(sd: SD, p: str, x):
    """text_grounding_net.py:75-81: Linear-SiLU-Linear-SiLU-Linear."""
    return linear(sd, p + ".4", F.silu(linear(sd, p + ".2", F.silu(linear(sd, p + ".0", x)))))


def unifusion(sd: SD, p: str, gi: Dict[str, torch.Tensor], flags: Dict[str, bool]):
    """UniFusion.forward in eval mode, text_grounding_net.py:185-313.  All-zero `segs` contribute exactly
    the null feature (:279-283); non-zero `segs` run the ConvNeXt mask encoder (:226-231).
    flags: test_drop_{boxes,points,scribbles,masks} of the config."""
    boxes, masks, text = gi["boxes"], gi["masks"], gi["positive_embeddings"]
    scribbles, polygons, segs, points = gi["scribbles"], gi["polygons"], gi["segs"], gi["points"]
    B, N, _ = boxes.shape
    m = masks.unsqueeze(-1)
    drop_box = flags.get("test_drop_boxes", False)
    drop_point = flags.get("test_drop_points", False)
    drop_scribble = flags.get("test_drop_scribbles", True)
    drop_polygons = flags.get("test_drop_masks", False)
    drop_segs = drop_polygons
    if drop_point and drop_box and drop_scribble and drop_polygons and drop_segs:
        drop_box = False
    if points is None:
        points = (boxes[:, :, :2] + boxes[:, :, 2:]) / 2.0
    text = text * m + (1 - m) * sd[p + ".null_positive_feature"].view(1, 1, -1)

    def sub(emb, msk, null):
        return emb * msk + (1 - msk) * sd[p + "." + null].view(1, 1, -1)

    zeros = torch.zeros_like(m)
    e_box = sub(fourier_embed(boxes), zeros if drop_box else m, "null_position_feature")
    e_pt = sub(fourier_embed(points), zeros if drop_point else m, "null_point_feature")
    m_s = zeros if drop_scribble else ((scribbles.sum(-1, keepdim=True) + m) > 0).float()
    e_s = sub(fourier_embed(scribbles), m_s, "null_scribble_feature")
    m_p = zeros if drop_polygons else ((polygons.sum(-1, keepdim=True) + m) > 0).float()
    e_p = sub(fourier_embed(polygons), m_p, "null_polygon_feature")
    seg_null = sd[p + ".null_seg_feature"].view(1, 1, -1).repeat(B, 64, 1)
    if p + ".in_conv.weight" in sd and not drop_segs and bool((segs.sum(dim=(1, 2, 3)) > 0).any()):
        # :226-231: resize, 30->3 conv, ConvNeXt-tiny, NCHW (B,768,16,16) reinterpreted as (B,3072,64) -> (B,64,3072)
        rs = F.interpolate(segs.float(), 512, mode="nearest")
        feat = convnext(sd, p + ".convnext_tiny_backbone", F.conv2d(rs, sd[p + ".in_conv.weight"], sd[p + ".in_conv.bias"], padding=1))
        feat = feat.reshape(B, -1, 64).permute(0, 2, 1)
        m_seg = (rs.sum(dim=(1, 2, 3)) > 0).float().view(-1, 1, 1)           # :279
        seg = feat * m_seg + (1 - m_seg) * seg_null + sd[p + ".pos_embedding"]  # :282-285
    else:
        # all-zero / dropped segs (or a state dict without the mask encoder): exactly the null feature
        if not drop_segs and bool((segs.sum(dim=(1, 2, 3)) > 0).any()):
            raise NotImplementedError("torch_oracle.unifusion: non-zero segs need the convnext / in_conv weights")
        seg = seg_null + sd[p + ".pos_embedding"]
    objs = [
        _mlp(sd, p + ".linears_list.0", torch.cat([text, e_box], -1)),
        _mlp(sd, p + ".linears_list.1", torch.cat([text, e_pt], -1)),
        _mlp(sd, p + ".linears_list.2", torch.cat([text, e_s], -1)),
        _mlp(sd, p + ".linears_list.3", torch.cat([text, e_p], -1)),
        _mlp(sd, p + ".linears_list.4", seg),
    ]
    return torch.cat(objs, dim=1), bool(drop_box and drop_polygons)


def null_grounding_input(gi: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """GroundingNetInput.get_null_input, text_grounding_tokinzer_input.py:56-93: zeros of the same shapes."""
    return {k: torch.zeros_like(v) if k != "segs" else v * 0 for k, v in gi.items()}


# ------------------------------------------------------------------------------------------------
# UNetModel.forward_single_input, openaimodel.py:482-563
# ------------------------------------------------------------------------------------------------
def _run_block(sd: SD, blk: str, h, emb, ctx, objs, scale):
    i = 0
    while True:
        q = f"{blk}.{i}"
        if q + ".in_layers.0.weight" in sd:
            h = resblock(sd, q, h, emb)
        elif q + ".proj_in.weight" in sd:
            h = spatial_transformer(sd, q, h, ctx, objs, scale)
        elif q + ".op.weight" in sd:
            h = downsample(sd, q, h)
        elif q + ".conv.weight" in sd:
            h = upsample(sd, q, h)
        elif q + ".weight" in sd and i == 0:  # input_blocks.0.0: the first conv
            h = F.conv2d(h, sd[q + ".weight"], sd[q + ".bias"], padding=1)
        else:
            break
        i += 1
    return h


def unet_forward(sd: SD, x, timesteps, context, gi: Dict[str, torch.Tensor], flags: Dict[str, bool],
                 scale: float = 1.0, first_conv: Optional[Dict[str, torch.Tensor]] = None):
    """gi: grounding dict (pass null_grounding_input(...) for the CFG-uncond branch, :483-487).
    scale: the fusers' `.scale` (set_alpha_scale).  first_conv: SD1.5 weights swapped in on alpha=0
    steps (restore_first_conv_from_SD, :469-480)."""
    if first_conv is not None:
        sd = dict(sd)
        sd["input_blocks.0.0.weight"] = first_conv["weight"]
        sd["input_blocks.0.0.bias"] = first_conv["bias"]
    objs, _ = unifusion(sd, "position_net", gi, flags)
    emb = linear(sd, "time_embed.2", F.silu(linear(sd, "time_embed.0", timestep_embedding(timesteps, 320))))
    h = x
    hs = []
    i = 0
    while f"input_blocks.{i}.0.weight" in sd or f"input_blocks.{i}.0.in_layers.0.weight" in sd \
            or f"input_blocks.{i}.0.op.weight" in sd:
        h = _run_block(sd, f"input_blocks.{i}", h, emb, context, objs, scale)
        hs.append(h)
        i += 1
    h = _run_block(sd, "middle_block", h, emb, context, objs, scale)
    i = 0
    while f"output_blocks.{i}.0.in_layers.0.weight" in sd:
        skip = hs.pop()
        b = torch.tanh(sd[f"scaleu_b_{i}"]) + 1
        s = torch.tanh(sd[f"scaleu_s_{i}"]) + 1
        h = torch.einsum("bchw,c->bchw", h, b)
        skip = fourier_filter(skip, 1, s)
        h = torch.cat([h, skip], dim=1)
        h = _run_block(sd, f"output_blocks.{i}", h, emb, context, objs, scale)
        i += 1
    h = F.silu(F.group_norm(h.float(), 32, sd["out.0.weight"], sd["out.0.bias"], 1e-5))
    return F.conv2d(h, sd["out.2.weight"], sd["out.2.bias"], padding=1)


# ------------------------------------------------------------------------------------------------
# samplers: ldm/models/diffusion/plms.py, plms_instance.py; utils/model.py:83-117
# ------------------------------------------------------------------------------------------------
def alphas_cumprod(linear_start=0.00085, linear_end=0.012, T=1000):
    """ddpm.py:21-36 + util.py:31-34 (fp64 -> fp32)."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, T, dtype=torch.float64) ** 2).numpy()
    return torch.tensor(np.cumprod(1. - betas, axis=0), dtype=torch.float32)


def alpha_schedule(length, alpha_type):
    n0 = int(alpha_type[0] * length)
    n1 = int(alpha_type[1] * length)
    decay = list(np.arange(start=0, stop=1, step=1 / n1)[::-1]) if n1 else []
    return [1] * n0 + decay + [0] * (length - n0 - n1)


def plms_sample(eval_fn: Callable, inputs: List[dict], uc, S: int, guidance: float, mis: float,
                alpha_type=None):
    """eval_fn(input_dict, alpha) -> eps.  inputs: [global, inst_1..] (one entry without MIS).
    Restates plms_instance.py:65-212 (== plms.py:72-167 when mis == 0)."""
    acp = alphas_cumprod()
    steps = np.asarray(list(range(0, 1000, 1000 // S))) + 1                  # util.py:55-70
    a = acp[steps]
    a_prev = torch.tensor([acp[0].item()] + acp[steps[:-1]].tolist(), dtype=torch.float32)
    time_range = np.flip(steps)
    total = len(steps)
    alphas = alpha_schedule(total, alpha_type) if alpha_type is not None else [1] * total
    mis_step = int(total * mis)

    def model_out(inp, alpha):
        e = eval_fn(inp, alpha)
        if uc is not None and guidance != 1:
            e_u = eval_fn(dict(x=inp["x"], timesteps=inp["timesteps"], context=uc), alpha)
            e = e_u + guidance * (e - e_u)
        return e

    def step(inp, old, i):
        index = total - i - 1
        b = inp["x"].shape[0]
        x = inp["x"].clone()
        dev = inp["x"].device
        t = torch.full((b,), int(time_range[i]), dtype=torch.long, device=dev)
        t_next = torch.full((b,), int(time_range[min(i + 1, total - 1)]), dtype=torch.long, device=dev)
        at, ap = a[index], a_prev[index]

        def x_prev_of(e):
            pred_x0 = (x - torch.sqrt(1. - at) * e) / at.sqrt()
            return ap.sqrt() * pred_x0 + (1. - ap).sqrt() * e

        inp["timesteps"] = t
        e_t = model_out(inp, alphas[i])
        if len(old) == 0:
            inp["x"] = x_prev_of(e_t)
            inp["timesteps"] = t_next
            e_p = (e_t + model_out(inp, alphas[i])) / 2
        elif len(old) == 1:
            e_p = (3 * e_t - old[-1]) / 2
        elif len(old) == 2:
            e_p = (23 * e_t - 16 * old[-1] + 5 * old[-2]) / 12
        else:
            e_p = (55 * e_t - 59 * old[-1] + 37 * old[-2] - 9 * old[-3]) / 24
        inp["x"] = x_prev_of(e_p)
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)

    olds = [[] for _ in inputs]
    for k, inp in enumerate(inputs):
        for i in range(mis_step):
            step(inp, olds[k], i)
    inputs[0]["x"] = torch.mean(torch.stack([inp["x"] for inp in inputs]), dim=0)
    for i in range(mis_step, total):
        step(inputs[0], olds[0], i)
    return inputs[0]["x"]


# ------------------------------------------------------------------------------------------------
# first-stage model: ldm/models/autoencoder.py, ldm/modules/diffusionmodules/model.py
# ------------------------------------------------------------------------------------------------
VAE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                    ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)  # configs/test_*.yaml:47-61
VAE_SCALE = 0.18215                                                                            # configs/test_*.yaml:45


def _gn6(sd: SD, p: str, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)  # model.py:38-39


def vae_resnet_block(sd: SD, p: str, x):
    """ResnetBlock.forward with temb=None, model.py:119-141."""
    h = F.conv2d(F.silu(_gn6(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn6(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".nin_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def vae_attn_block(sd: SD, p: str, x):
    """AttnBlock.forward, model.py:177-202."""
    h = _gn6(sd, p + ".norm", x)
    q = F.conv2d(h, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    w_ = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
    w_ = torch.softmax(w_, dim=2)
    h = torch.bmm(v.reshape(b, c, hh * ww), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + F.conv2d(h, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def vae_decode(sd: SD, z, cfg: dict = VAE_DDCONFIG, scale: float = VAE_SCALE):
    """AutoencoderKL.decode (autoencoder.py:33-37) + Decoder.forward (model.py:528-569)."""
    nres, nb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    h = F.conv2d(z.float() / scale, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    d = "decoder"
    h = F.conv2d(h, sd[d + ".conv_in.weight"], sd[d + ".conv_in.bias"], padding=1)
    h = vae_resnet_block(sd, d + ".mid.block_1", h)
    h = vae_attn_block(sd, d + ".mid.attn_1", h)
    h = vae_resnet_block(sd, d + ".mid.block_2", h)
    for lvl in reversed(range(nres)):
        for i in range(nb + 1):
            h = vae_resnet_block(sd, f"{d}.up.{lvl}.block.{i}", h)
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[f"{d}.up.{lvl}.upsample.conv.weight"], sd[f"{d}.up.{lvl}.upsample.conv.bias"], padding=1)
    h = F.silu(_gn6(sd, d + ".norm_out", h))
    return F.conv2d(h, sd[d + ".conv_out.weight"], sd[d + ".conv_out.bias"], padding=1)


def vae_encode_moments(sd: SD, x, cfg: dict = VAE_DDCONFIG):
    """quant_conv(Encoder.forward(x)) (autoencoder.py:27-29, model.py:429-459): the (mean | logvar) planes the
    posterior is sampled from (the sample itself draws torch.randn, so parity is stated on the moments)."""
    nres, nb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    e = "encoder"
    h = F.conv2d(x.float(), sd[e + ".conv_in.weight"], sd[e + ".conv_in.bias"], padding=1)
    for lvl in range(nres):
        for i in range(nb):
            h = vae_resnet_block(sd, f"{e}.down.{lvl}.block.{i}", h)
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)  # model.py:70-73
            h = F.conv2d(h, sd[f"{e}.down.{lvl}.downsample.conv.weight"], sd[f"{e}.down.{lvl}.downsample.conv.bias"], stride=2)
    h = vae_resnet_block(sd, e + ".mid.block_1", h)
    h = vae_attn_block(sd, e + ".mid.attn_1", h)
    h = vae_resnet_block(sd, e + ".mid.block_2", h)
    h = F.conv2d(F.silu(_gn6(sd, e + ".norm_out", h)), sd[e + ".conv_out.weight"], sd[e + ".conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


# ------------------------------------------------------------------------------------------------
# CLIP text encoder.  Third-party arithmetic: Hugging Face transformers' CLIPTextModel,
# pinned transformers==4.27.0 by the reference (requirements.txt:247), called at ldm/modules/encoders/modules.py:
# 147-165 (FrozenCLIPEmbedder) and utils/model.py:146-151 (get_clip_feature).  Published algorithm
# (modeling_clip.py: CLIPTextEmbeddings, CLIPEncoderLayer, CLIPAttention, CLIPMLP with quick_gelu,
# CLIPTextTransformer), restated over the model's own state_dict; pinned by tests/golden/clip_text.pt, produced by
# the installed transformers CLIPTextModel itself (oracle/make_golden.py --only clip).
# ------------------------------------------------------------------------------------------------
CLIP_TEXT_CONFIG = dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                        num_attention_heads=12, max_position_embeddings=77)


def clip_text_forward(sd: SD, input_ids: torch.Tensor, heads: int = 12, eps: float = 1e-5, prefix: str = "text_model",
                      key_len: Optional[torch.Tensor] = None):
    """-> (last_hidden_state (B, T, C), pooler_output (B, C)).  Pre-LN transformer, causal mask (plus key padding when
    key_len is given), QuickGELU x * sigmoid(1.702 x), final LayerNorm, pooled = state at the highest token id (EOS)."""
    B, T = input_ids.shape
    x = sd[f"{prefix}.embeddings.token_embedding.weight"][input_ids] + sd[f"{prefix}.embeddings.position_embedding.weight"][:T][None]
    C = x.shape[-1]
    d = C // heads
    mask = torch.full((T, T), float("-inf"), device=x.device).triu(1)[None, None]  # key j visible to query i iff j <= i
    if key_len is not None:
        pad = torch.arange(T, device=x.device)[None, :] >= key_len.to(x.device)[:, None]
        mask = mask + torch.zeros((B, 1, T, T), device=x.device).masked_fill(pad[:, None, None, :], float("-inf"))
    n_layers = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(f"{prefix}.encoder.layers."))
    for i in range(n_layers):
        L = f"{prefix}.encoder.layers.{i}"
        h = F.layer_norm(x, (C,), sd[L + ".layer_norm1.weight"], sd[L + ".layer_norm1.bias"], eps)
        q = F.linear(h, sd[L + ".self_attn.q_proj.weight"], sd[L + ".self_attn.q_proj.bias"]) * d ** -0.5
        k = F.linear(h, sd[L + ".self_attn.k_proj.weight"], sd[L + ".self_attn.k_proj.bias"])
        v = F.linear(h, sd[L + ".self_attn.v_proj.weight"], sd[L + ".self_attn.v_proj.bias"])
        sp = lambda t: t.view(B, T, heads, d).transpose(1, 2)
        a = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) + mask, dim=-1) @ sp(v)
        a = a.transpose(1, 2).reshape(B, T, C)
        x = x + F.linear(a, sd[L + ".self_attn.out_proj.weight"], sd[L + ".self_attn.out_proj.bias"])
        h = F.layer_norm(x, (C,), sd[L + ".layer_norm2.weight"], sd[L + ".layer_norm2.bias"], eps)
        u = F.linear(h, sd[L + ".mlp.fc1.weight"], sd[L + ".mlp.fc1.bias"])
        x = x + F.linear(u * torch.sigmoid(1.702 * u), sd[L + ".mlp.fc2.weight"], sd[L + ".mlp.fc2.bias"])
    last = F.layer_norm(x, (C,), sd[f"{prefix}.final_layer_norm.weight"], sd[f"{prefix}.final_layer_norm.bias"], eps)
    pooled = last[torch.arange(B, device=x.device), input_ids.argmax(dim=-1)]
    return last, pooled

