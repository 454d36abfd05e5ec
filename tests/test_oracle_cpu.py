"""CPU suite, part 1: pin the plain-torch restatement (oracle/torch_oracle.py) against the golden
fixtures that the reference's own modules produced (oracle/make_golden.py).  fp32 vs fp32 on the
same seeded inputs: only summation order differs, so the tolerances are tight."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
from oracle import torch_oracle as TO  # noqa: E402
from instancediffusion_b200 import synthetic  # noqa: E402
from instancediffusion_b200.weights import UNIFUSION_FLAGS, synth_tensor  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    return torch.load(path, map_location="cpu")


def _close(got, ref, rtol, atol, what):
    err = (got - ref).abs().max().item()
    assert torch.allclose(got, ref, rtol=rtol, atol=atol), f"{what}: max abs err {err:.3e}"


def _case_sd(name, keys_shapes):
    return {k: synth_tensor(f"{name}.{k}", s, cases.WEIGHT_SEED) for k, s in keys_shapes.items()}


def _module_shapes(spec):
    """state_dict shapes of the case's module, from the mirror class built on the meta device."""
    import importlib
    mod, cls = spec["module"].split(":")
    klass = getattr(importlib.import_module("instancediffusion_b200.ldm.modules." + mod), cls)
    with torch.device("meta"):
        m = klass(*spec["args"], **spec.get("kwargs", {}))
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


_RUNNERS = {
    "GatedSelfAttentionDense": lambda sd, i: TO.gated_self_attention(sd, "m", i["x"], i["objs"]),
    "SelfAttention": lambda sd, i: TO.self_attention(sd, "m", i["x"]),
    "CrossAttention": lambda sd, i: TO.cross_attention(sd, "m", i["x"], i["key"]),
    "FeedForward": lambda sd, i: TO.feed_forward(sd, "m", i["x"]),
    "BasicTransformerBlock": lambda sd, i: TO.basic_transformer_block(sd, "m", i["x"], i["context"], i["objs"]),
    "SpatialTransformer": lambda sd, i: TO.spatial_transformer(sd, "m", i["x"], i["context"], i["objs"]),
    "ResBlock": lambda sd, i: TO.resblock(sd, "m", i["x"], i["emb"]),
    "Upsample": lambda sd, i: TO.upsample(sd, "m", i["x"]),
    "Downsample": lambda sd, i: TO.downsample(sd, "m", i["x"]),
}


@pytest.mark.parametrize("name", list(cases.MODULE_CASES))
def test_module_restatement_matches_reference(name):
    gold = _load("modules.pt")
    spec = cases.MODULE_CASES[name]
    sd = {"m." + k: v for k, v in _case_sd(name, _module_shapes(spec)).items()}
    ins = cases.build_inputs(name, spec)
    with torch.no_grad():
        out = _RUNNERS[spec["module"].split(":")[1]](sd, ins)
    _close(out, gold[name], 1e-4, 2e-5, name)


def test_fourier_filter_and_timestep_embedding():
    gold = _load("fourier.pt")
    for name, spec in cases.FOURIER_CASES.items():
        x = cases.synth_input(name, "x", spec["shape"]) + 0.5
        _close(TO.fourier_filter(x, 1, spec["scale"]), gold[name], 1e-5, 1e-5, name)
    t = torch.tensor([981, 1, 501, 21])
    _close(TO.timestep_embedding(t, 320), gold["timestep_embedding"], 0, 1e-6, "timestep_embedding")


def test_fourier_filter_closed_form():
    """The identity the ScaleU kernel implements: filter(x) = x + (s-1) * P_low(x) with P_low the
    real part of the inverse DFT over bins {-1,0}^2 (7 real sums per plane)."""
    import math
    for name, spec in cases.FOURIER_CASES.items():
        x = (cases.synth_input(name, "x", spec["shape"]) + 0.5).double()
        B, C, H, W = x.shape
        yy = torch.arange(H, dtype=torch.float64)[:, None] * (2 * math.pi / H)
        xx = torch.arange(W, dtype=torch.float64)[None, :] * (2 * math.pi / W)
        basis = [torch.ones(H, W, dtype=torch.float64), torch.cos(xx).expand(H, W), torch.sin(xx).expand(H, W),
                 torch.cos(yy).expand(H, W), torch.sin(yy).expand(H, W), torch.cos(xx + yy), torch.sin(xx + yy)]
        plow = sum((x * b).sum((-2, -1), keepdim=True) * b for b in basis) / (H * W)
        closed = x + (spec["scale"] - 1) * plow
        ref = TO.fourier_filter(x.float(), 1, spec["scale"]).double()
        assert (closed - ref).abs().max() < 5e-6, name


@pytest.mark.parametrize("name", list(cases.UNIFUSION_CASES))
def test_unifusion_restatement_matches_reference(name):
    gold = _load("unifusion.pt")
    spec = cases.UNIFUSION_CASES[name]
    from instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net import UniFusion
    with torch.device("meta"):
        net = UniFusion(in_dim=768, out_dim=768, mid_dim=3072)
    sd = {"position_net." + k: synth_tensor("position_net." + k, tuple(v.shape), 0)
          for k, v in net.state_dict().items() if "convnext" not in k and "in_conv" not in k}
    gb = synthetic.make_grounding_batch(spec["batch"], spec["n"], spec["seed"], spec["flavor"])
    gi = dict(boxes=gb["boxes"], masks=gb["masks"], positive_embeddings=gb["text_embeddings"],
              scribbles=gb["scribbles"], polygons=gb["polygons"], segs=gb["segs"], points=gb["points"])
    with torch.no_grad():
        objs, dbm = TO.unifusion(sd, "position_net", gi, UNIFUSION_FLAGS[spec["flavor"]])
    _close(objs, gold[name], 1e-4, 2e-5, name)
    assert int(dbm) == int(gold[name + "/drop_box_mask"])


@pytest.mark.parametrize("name", list(cases.CONVNEXT_CASES))
def test_convnext_restatement_matches_reference(name):
    gold = _load("convnext.pt")
    spec = cases.CONVNEXT_CASES[name]
    from instancediffusion_b200.ldm.modules.diffusionmodules import convnext as cnx
    with torch.device("meta"):
        m = getattr(cnx, spec["cls"])(*spec["args"])
    sd = {"m." + k: synth_tensor(f"{name}.{k}", tuple(v.shape), cases.WEIGHT_SEED) for k, v in m.state_dict().items()}
    x = cases.synth_input(name, "x", spec["inputs"]["x"])
    with torch.no_grad():
        out = TO.convnext_block(sd, "m", x) if spec["cls"] == "Block" else TO.convnext(sd, "m", x)
    _close(out, gold[name], 1e-4, 2e-5, name)


@pytest.mark.parametrize("name", list(cases.UNIFUSION_MASK_CASES))
def test_unifusion_mask_restatement_matches_reference(name):
    """Non-zero `segs`: in_conv + ConvNeXt + token reinterpretation of the oracle port vs the reference."""
    gold = _load("unifusion_mask.pt")
    spec = cases.UNIFUSION_MASK_CASES[name]
    from instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net import UniFusion
    with torch.device("meta"):
        net = UniFusion(in_dim=768, out_dim=768, mid_dim=3072)
    sd = {"position_net." + k: synth_tensor("position_net." + k, tuple(v.shape), 0) for k, v in net.state_dict().items()}
    gb = synthetic.make_grounding_batch(spec["batch"], spec["n"], spec["seed"], spec["flavor"])
    gi = dict(boxes=gb["boxes"], masks=gb["masks"], positive_embeddings=gb["text_embeddings"],
              scribbles=gb["scribbles"], polygons=gb["polygons"], segs=gb["segs"], points=gb["points"])
    with torch.no_grad():
        objs, dbm = TO.unifusion(sd, "position_net", gi, UNIFUSION_FLAGS[spec["flavor"]])
    _close(objs, gold[name], 1e-4, 2e-5, name)
    assert int(dbm) == int(gold[name + "/drop_box_mask"])


def test_schema_matches_reference():
    """The mirror UNetModel exposes exactly the reference's 1199 state_dict keys and shapes
    (utils/checkpoint.py:241-244 loads strict)."""
    import json
    path = os.path.join(GOLDEN, "unet_schema.json")
    if not os.path.exists(path):
        pytest.skip("unet_schema.json not generated")
    ref_schema = json.load(open(path))
    from instancediffusion_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    from instancediffusion_b200.weights import unet_config
    with torch.device("meta"):
        m = UNetModel(**unet_config("box"))
    ours = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert len(ref_schema) == 1199
    assert set(ours) == set(ref_schema), (sorted(set(ours) ^ set(ref_schema))[:10])
    bad = [k for k in ours if ours[k] != ref_schema[k]]
    assert not bad, bad[:10]
    assert list(ours) == list(ref_schema), "key order differs"


# ------------------------------------------------------------------------------------------------
# whole-UNet and sampler restatements (what bench.py's CPU arm and smoke() execute) pinned to the
# reference's own outputs
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def unet_sd():
    import json
    path = os.path.join(GOLDEN, "unet_schema.json")
    if not os.path.exists(path):
        pytest.skip("unet_schema.json not generated")
    schema = json.load(open(path))
    return {k: synth_tensor(k, tuple(s), cases.UNET_CASE["weight_seed"]) for k, s in schema.items()}


def _rel_l2(got, ref):
    return ((got - ref).norm() / ref.norm()).item()


def _gti():
    from instancediffusion_b200.grounding_input.text_grounding_tokinzer_input import GroundingNetInput
    return GroundingNetInput()


def test_unet_forward_restatement_matches_reference(unet_sd):
    """oracle/torch_oracle.py:unet_forward vs the reference's UNetModel.forward (tests/golden/unet.pt):
    fp32 vs fp32, only the summation order of a few fused expressions differs."""
    gold = _load("unet.pt")
    spec = cases.UNET_CASE
    inp, uc = synthetic.make_sampler_inputs(_gti(), spec["batch"], spec["n"], spec["seed"], spec["flavor"], mis=False)
    ts = torch.full((spec["batch"],), spec["t"], dtype=torch.long)
    gi = inp["grounding_input"]
    flags = UNIFUSION_FLAGS[spec["flavor"]]
    sd15 = _load("sd15_first_conv.pt")
    with torch.no_grad():
        e_c = TO.unet_forward(unet_sd, inp["x"], ts, inp["context"], gi, flags)
        e_u = TO.unet_forward(unet_sd, inp["x"], ts, uc, TO.null_grounding_input(gi), flags)
        e_0 = TO.unet_forward(unet_sd, inp["x"], ts, inp["context"], gi, flags, scale=0.0, first_conv=sd15)
    for got, key in ((e_c, "eps_cond"), (e_u, "eps_null"), (e_0, "eps_alpha0")):
        rel = _rel_l2(got, gold[key])
        print(f"[oracle unet_forward] {key}: rel_l2 {rel:.2e}")
        assert rel < 1e-4, (key, rel)
        _close(got, gold[key], 1e-3, 1e-4, key)  # north_star's latent tolerance, met fp32-vs-fp32


@pytest.mark.parametrize("name", list(cases.SAMPLER_CASES))
def test_plms_sample_restatement_matches_reference(unet_sd, name):
    """oracle/torch_oracle.py:plms_sample (PLMS and the Multi-instance Sampler) vs the latents the
    reference's PLMSSampler / PLMSSamplerInst produced (tests/golden/samplers.pt)."""
    gold = _load("samplers.pt")
    sc = cases.SAMPLER_CASES[name]
    sd15 = _load("sd15_first_conv.pt")
    flags = UNIFUSION_FLAGS["box"]
    inputs, uc = synthetic.make_sampler_inputs(_gti(), sc["batch"], sc["n"], sc["seed"], "box", mis=sc["mis"] > 0)
    inputs = inputs if isinstance(inputs, list) else [inputs]
    null_gi = TO.null_grounding_input(inputs[0]["grounding_input"])

    def eval_fn(inp, alpha):
        gi = inp.get("grounding_input")
        return TO.unet_forward(unet_sd, inp["x"], inp["timesteps"], inp["context"], gi if gi is not None else null_gi,
                               flags, scale=float(alpha), first_conv=sd15 if alpha == 0 else None)

    with torch.no_grad():
        x = TO.plms_sample(eval_fn, inputs, uc, sc["S"], sc["guidance"], sc["mis"], alpha_type=sc["alpha_type"])
    rel = _rel_l2(x, gold[name])
    print(f"[oracle plms_sample] {name}: rel_l2 {rel:.2e}")
    assert rel < 2e-4, rel


# ------------------------------------------------------------------------------------------------
# first-stage model (AutoencoderKL.decode / encoder moments)
# ------------------------------------------------------------------------------------------------
def _vae_sd():
    from instancediffusion_b200.ldm.models.autoencoder import AutoencoderKL
    with torch.device("meta"):
        m = AutoencoderKL(dict(TO.VAE_DDCONFIG), 4, TO.VAE_SCALE)
    return {k: synth_tensor("vae." + k, tuple(v.shape), cases.WEIGHT_SEED) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", list(cases.VAE_CASES))
def test_vae_restatement_matches_reference(name):
    """oracle vae_decode / vae_encode_moments vs the reference's AutoencoderKL (tests/golden/vae.pt)."""
    gold = _load("vae.pt")
    spec = cases.VAE_CASES[name]
    sd = _vae_sd()
    g = torch.Generator().manual_seed(spec["seed"])
    with torch.no_grad():
        if spec["kind"] == "decode":
            z = torch.randn((spec["batch"], 4, spec["size"], spec["size"]), generator=g) * spec["std"]
            got = TO.vae_decode(sd, z)
        else:
            x = torch.randn((spec["batch"], 3, spec["size"], spec["size"]), generator=g) * spec["std"]
            got = TO.vae_encode_moments(sd, x)
    _close(got, gold[name], 1e-4, 5e-5, f"vae {name}")


# ------------------------------------------------------------------------------------------------
# instance-isolation attention mask (efficient_attention=False path of the gated self-attention)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.MASKED_CASES))
def test_masked_gated_attention_restatement_matches_reference(name):
    gold = _load("masked.pt")
    spec = cases.MASKED_CASES[name]
    x, objs, boxes, counts, att = cases.masked_case_inputs(name, spec)
    from instancediffusion_b200.ldm.modules.attention import GatedSelfAttentionDense
    with torch.device("meta"):
        m = GatedSelfAttentionDense(*spec["args"], efficient_attention=False)
    sd = {"m." + k: synth_tensor(f"{name}.{k}", tuple(v.shape), cases.WEIGHT_SEED) for k, v in m.state_dict().items()}
    with torch.no_grad():
        got = TO.gated_self_attention(sd, "m", x, objs, att_masks=att)
        free = TO.gated_self_attention(sd, "m", x, objs)
    _close(got[:, ::spec["stride"]], gold[name], 1e-4, 2e-5, name)
    _close(free[:, ::spec["stride"]], gold[name + "/free"], 1e-4, 2e-5, name + "/free")


# ------------------------------------------------------------------------------------------------
# CLIP text tower (third-party: transformers CLIPTextModel; host prep)
# ------------------------------------------------------------------------------------------------
def _clip_sd():
    from instancediffusion_b200.ldm.modules.encoders.modules import CLIPTextModel
    from instancediffusion_b200.weights import synth_tensor
    with torch.device("meta"):
        m = CLIPTextModel()
    return {k: synth_tensor("clip." + k, tuple(v.shape), cases.WEIGHT_SEED) for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", list(cases.CLIP_CASES))
def test_clip_text_restatement_matches_transformers(name):
    """oracle clip_text_forward vs the golden produced by the installed transformers CLIPTextModel with the same
    synthetic weights (tests/golden/clip_text.pt); the mirror's state_dict keys are HF's, so the same dict feeds both."""
    gold = _load("clip_text.pt")
    ids = cases.clip_token_ids(cases.CLIP_CASES[name])
    with torch.no_grad():
        last, pooled = TO.clip_text_forward(_clip_sd(), ids)
    _close(last, gold[name + "/last_hidden_state"], 1e-4, 5e-5, f"clip {name} last_hidden_state")
    _close(pooled, gold[name + "/pooler_output"], 1e-4, 5e-5, f"clip {name} pooler_output")


def test_clip_text_key_padding_mask_matches_transformers():
    """get_clip_feature (utils/model.py:146-151) passes the processor's attention_mask: the restated key-padding mask
    against transformers on a padded batch (tiny random model built here: no fixture needed)."""
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(3)
    cfg = CLIPTextConfig(vocab_size=100, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         max_position_embeddings=16, hidden_act="quick_gelu", eos_token_id=99, bos_token_id=98, pad_token_id=0)
    m = CLIPTextModel(cfg).eval()
    ids = torch.tensor([[98, 5, 6, 99, 0, 0, 0, 0], [98, 7, 8, 9, 10, 11, 12, 99]])
    am = (torch.arange(8)[None] < torch.tensor([4, 8])[:, None]).long()
    with torch.no_grad():
        ref = m(input_ids=ids, attention_mask=am)
        last, pooled = TO.clip_text_forward(m.state_dict(), ids, heads=4, key_len=am.sum(-1))
    # rows of the padding region depend on the (implementation-defined) treatment of fully masked queries: compare real tokens
    for b, n in enumerate((4, 8)):
        _close(last[b, :n], ref.last_hidden_state[b, :n], 1e-4, 5e-5, f"clip padded batch row {b}")
    _close(pooled, ref.pooler_output, 1e-4, 5e-5, "clip padded batch pooled")
