"""`-m gpu` parity tests of the drop-in modules against the golden fixtures produced by the
reference's own modules (CPU fp32, oracle/make_golden.py).  The CUDA path computes in fp16 with
fp32 accumulation -- what the reference does under torch.autocast(fp16) (inference.py:94) -- so the
tolerance is an fp16 one, stated per test: relative L2 error and max abs error relative to the
reference's dynamic range.
"""
import importlib
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(HERE, "golden")


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not generated")
    return torch.load(path, map_location="cpu")


def _report(got, ref, what, rel_l2_tol, max_tol):
    got = got.detach().float().cpu()
    ref = ref.float()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    rel = ((got - ref).norm() / ref.norm()).item()
    mx = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"[{what}] rel_l2={rel:.3e} max_err/ref_max={mx:.3e} (tol {rel_l2_tol:.0e} / {max_tol:.0e})")
    assert rel < rel_l2_tol and mx < max_tol, f"{what}: rel_l2 {rel:.3e}, max {mx:.3e}"


def _mirror_class(path):
    mod, cls = path.split(":")
    return getattr(importlib.import_module("instancediffusion_b200.ldm.modules." + mod), cls)


@pytest.mark.parametrize("name", list(cases.MODULE_CASES))
def test_module_matches_reference_golden(cuda_device, name):
    gold = _load("modules.pt")
    spec = cases.MODULE_CASES[name]
    out = cases.run_module_case(name, spec, _mirror_class(spec["module"]), device=cuda_device)
    # fp16 operands / fp32 accumulation through <= ~12 chained GEMMs: 3e-3 relative L2
    _report(out, gold[name], name, 1.4e-3, 2.1e-3)  # <= 2x measured (max over the cases: 7.0e-4 / 1.07e-3)


def test_fourier_filter_matches_reference_golden(cuda_device):
    from instancediffusion_b200.ldm.modules.diffusionmodules.openaimodel import Fourier_filter
    gold = _load("fourier.pt")
    for name, spec in cases.FOURIER_CASES.items():
        x = (cases.synth_input(name, "x", spec["shape"]) + 0.5).to(cuda_device)
        out = Fourier_filter(x, threshold=1, scale=spec["scale"])
        _report(out, gold[name], name, 6e-4, 1.2e-3)  # fp16 in/out of an fp32 closed form; measured 3.1e-4 / 6.3e-4
    from instancediffusion_b200 import ops
    t = torch.tensor([981.0, 1.0, 501.0, 21.0], device=cuda_device)
    _report(ops.timestep_embedding(t, 320), gold["timestep_embedding"], "timestep_embedding", 3e-4, 5e-4)  # measured 1.5e-4 / 2.4e-4


@pytest.mark.parametrize("name", list(cases.UNIFUSION_CASES))
def test_unifusion_matches_reference_golden(cuda_device, name):
    from instancediffusion_b200 import synthetic
    from instancediffusion_b200.grounding_input.text_grounding_tokinzer_input import GroundingNetInput
    from instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net import UniFusion
    from instancediffusion_b200.weights import UNIFUSION_FLAGS, load_synthetic
    gold = _load("unifusion.pt")
    spec = cases.UNIFUSION_CASES[name]
    with torch.device("meta"):
        net = UniFusion(in_dim=768, out_dim=768, mid_dim=3072, **UNIFUSION_FLAGS[spec["flavor"]])
    net = net.to_empty(device=cuda_device).eval()
    load_synthetic(net, 0, prefix="position_net.")
    gb = synthetic.make_grounding_batch(spec["batch"], spec["n"], spec["seed"], spec["flavor"], device=cuda_device)
    gi = GroundingNetInput().prepare(gb)
    objs, dbm = net(gi["boxes"], gi["masks"], gi["positive_embeddings"], gi["scribbles"], gi["polygons"],
                    gi["segs"], gi["points"])
    _report(objs, gold[name], name, 1.1e-3, 1.2e-3)  # measured 5.6e-4 / 6.0e-4
    assert int(dbm) == int(gold[name + "/drop_box_mask"])


@pytest.mark.parametrize("name", list(cases.CONVNEXT_CASES))
def test_convnext_matches_reference_golden(cuda_device, name):
    """ConvNeXt Block / the whole ConvNeXt-tiny trunk vs the reference's convnext.py (CPU fp32)."""
    from instancediffusion_b200.ldm.modules.diffusionmodules import convnext as cnx
    from instancediffusion_b200.weights import load_synthetic
    gold = _load("convnext.pt")
    spec = cases.CONVNEXT_CASES[name]
    m = getattr(cnx, spec["cls"])(*spec["args"])
    load_synthetic(m, cases.WEIGHT_SEED, prefix=name + ".")
    m = m.to(cuda_device).eval()
    x = cases.synth_input(name, "x", spec["inputs"]["x"]).to(cuda_device)
    with torch.no_grad():
        out = m(x)
    # a Block is 2 GEMMs deep (3e-3 like the other modules); the trunk chains 18 blocks + 4 strided convs
    _report(out, gold[name], name, 6e-4 if "block" in name else 2.3e-3, 1.4e-3 if "block" in name else 2.2e-3)  # measured 3.0e-4 / 7.1e-4 (blocks), 1.16e-3 / 1.11e-3 (whole encoder)


@pytest.mark.parametrize("name", list(cases.UNIFUSION_MASK_CASES))
def test_unifusion_mask_matches_reference_golden(cuda_device, name):
    """Mask conditioning: non-zero `segs` through in_conv + ConvNeXt + the 64 mask tokens
    (text_grounding_net.py:226-231, 277-287), polygons live."""
    from instancediffusion_b200 import synthetic
    from instancediffusion_b200.grounding_input.text_grounding_tokinzer_input import GroundingNetInput
    from instancediffusion_b200.ldm.modules.diffusionmodules.text_grounding_net import UniFusion
    from instancediffusion_b200.weights import UNIFUSION_FLAGS, load_synthetic
    gold = _load("unifusion_mask.pt")
    spec = cases.UNIFUSION_MASK_CASES[name]
    with torch.device("meta"):
        net = UniFusion(in_dim=768, out_dim=768, mid_dim=3072, **UNIFUSION_FLAGS[spec["flavor"]])
    net = net.to_empty(device=cuda_device).eval()
    load_synthetic(net, 0, prefix="position_net.")
    gb = synthetic.make_grounding_batch(spec["batch"], spec["n"], spec["seed"], spec["flavor"], device=cuda_device)
    gi = GroundingNetInput().prepare(gb)
    # the ConvNeXt feature map on its own first (localises a failure)
    y, seg_sum = __import__("instancediffusion_b200.ops", fromlist=["x"]).segs_inconv(
        gi["segs"].float(), net.pk()["w_inconv"], net.pk()["b_inconv"], 512)
    feat, fh, fw = net.convnext_tiny_backbone._features(y, spec["batch"], 512, 512)
    ref_feat = gold[name + "/convnext_feat"]
    _report(feat.view(spec["batch"], fh, fw, -1).permute(0, 3, 1, 2), ref_feat, name + "/convnext_feat", 2.4e-3, 3e-3)  # measured 1.21e-3 / 1.50e-3
    objs, dbm = net(gi["boxes"], gi["masks"], gi["positive_embeddings"], gi["scribbles"], gi["polygons"],
                    gi["segs"], gi["points"])
    _report(objs, gold[name], name, 1.8e-3, 2.4e-3)  # measured 9.1e-4 / 1.21e-3
    assert int(dbm) == int(gold[name + "/drop_box_mask"])


# --------------------------------------------------------------------------------------------
# whole UNet + samplers (synthetic weights regenerated bit-identically from the seed)
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def unet(cuda_device):
    from instancediffusion_b200.weights import build_unet
    spec = cases.UNET_CASE
    model = build_unet(spec["flavor"], cuda_device, seed=spec["weight_seed"])
    sd_conv = _load("sd15_first_conv.pt")
    model._sd_conv = sd_conv
    return model


def _unet_inputs(model, device):
    from instancediffusion_b200 import synthetic
    spec = cases.UNET_CASE
    inp, uc = synthetic.make_sampler_inputs(model.grounding_tokenizer_input, spec["batch"], spec["n"], spec["seed"],
                                            spec["flavor"], mis=False, device=device)
    ts = torch.full((spec["batch"],), spec["t"], dtype=torch.long, device=device)
    return inp, uc, ts


def test_unet_eps_matches_reference_golden(cuda_device, unet):
    from instancediffusion_b200.utils.model import set_alpha_scale
    gold = _load("unet.pt")
    inp, uc, ts = _unet_inputs(unet, cuda_device)
    gi = inp["grounding_input"]
    set_alpha_scale(unet, 1)
    objs, _ = unet.position_net(gi["boxes"], gi["masks"], gi["positive_embeddings"], gi["scribbles"], gi["polygons"],
                                gi["segs"], gi["points"])
    _report(objs, gold["objs"], "unet/objs", 1.1e-3, 1.2e-3)
    # one full denoise forward: ~200 fp16 layers deep.  Measured 2.0e-3 relative L2;
    # the bound is 2x that.  (The reference itself under autocast(fp16) deviates as much:
    # tests/test_parity_r2_gpu.py::test_fp16_envelope_eps_and_latents.)
    for graph in (False, True):
        unet.use_cuda_graph = graph
        eps_c = unet(dict(x=inp["x"], timesteps=ts, context=inp["context"], grounding_input=gi))
        _report(eps_c, gold["eps_cond"], f"unet/eps_cond graph={graph}", 4e-3, 4.5e-3)
        eps_u = unet(dict(x=inp["x"], timesteps=ts, context=uc))
        _report(eps_u, gold["eps_null"], f"unet/eps_null graph={graph}", 4e-3, 4.5e-3)
    # batched cond+uncond: every row against the *reference golden* (not against our own single path), same
    # bound.  Tile widths / stream-K splits depend on M, so batched and single runs round differently at the
    # fp16 level and are not bit-equal; both must sit inside the same distance of the fp32 reference.
    both = unet.forward_batched([dict(x=inp["x"], timesteps=ts, context=inp["context"], grounding_input=gi),
                                 dict(x=inp["x"], timesteps=ts, context=uc)])
    _report(both[0], gold["eps_cond"], "batched cond vs golden", 4e-3, 4.5e-3)
    _report(both[1], gold["eps_null"], "batched uncond vs golden", 4e-3, 4.5e-3)
    _report(both[0], eps_c.cpu(), "batched cond vs single", 3.4e-3, 3.5e-3)  # measured 1.6-1.7e-3: two fp16 roundings of the same eps
    _report(both[1], eps_u.cpu(), "batched uncond vs single", 3.4e-3, 3.5e-3)
    # alpha = 0: fusers off + SD1.5 first conv (openaimodel.py:469-480)
    set_alpha_scale(unet, 0)
    unet.set_sd_first_conv(unet._sd_conv)
    eps_0 = unet(dict(x=inp["x"], timesteps=ts, context=inp["context"], grounding_input=gi))
    _report(eps_0, gold["eps_alpha0"], "unet/eps_alpha0", 4e-3, 4.5e-3)
    unet.undo_first_conv_restore()
    set_alpha_scale(unet, 1)


@pytest.mark.parametrize("name", list(cases.SAMPLER_CASES))
def test_sampler_latent_vs_reference_golden(cuda_device, unet, name):
    """End-to-end latent after the full PLMS / Multi-instance loop (config 1 of BASELINE.json for
    mis_S10).  north_star's rtol=1e-3/atol=1e-4 on the latent is tighter than fp16 re-association
    noise compounded over 10-30 CFG-7.5 forwards (measured for the reference's own arithmetic under
    autocast(fp16) in tests/test_parity_r2_gpu.py); the bound held here is 7.4e-3 relative L2 vs the fp32
    reference = 2x the measured value, which is printed."""
    from functools import partial
    from instancediffusion_b200 import synthetic
    from instancediffusion_b200.ldm.models.diffusion.ldm import LatentDiffusion
    from instancediffusion_b200.ldm.models.diffusion.plms import PLMSSampler
    from instancediffusion_b200.ldm.models.diffusion.plms_instance import PLMSSamplerInst
    from instancediffusion_b200.utils.model import alpha_generator, set_alpha_scale
    gold = _load("samplers.pt")
    sc = cases.SAMPLER_CASES[name]
    unet.undo_first_conv_restore()
    unet.use_cuda_graph = True
    os.environ["IDIFF_PRETRAINED_DIR"] = ""  # force the explicit path below
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(cuda_device)
    agen = partial(alpha_generator, type=sc["alpha_type"])
    use_mis = sc["mis"] > 0
    inputs, uc = synthetic.make_sampler_inputs(unet.grounding_tokenizer_input, sc["batch"], sc["n"], sc["seed"], "box",
                                               mis=use_mis, device=cuda_device)
    # the reference reads pretrained/SD_v1_5_input_conv_weight_bias.pth from the cwd at alpha == 0;
    # the same tensors are committed as a fixture
    orig = unet.restore_first_conv_from_SD
    unet.restore_first_conv_from_SD = lambda: (None if getattr(unet, "_first_conv_restored", False)
                                               else unet.set_sd_first_conv(unet._sd_conv))
    try:
        if use_mis:
            sampler = PLMSSamplerInst(diffusion, unet, alpha_generator_func=agen, set_alpha_scale=set_alpha_scale, mis=sc["mis"])
        else:
            sampler = PLMSSampler(diffusion, unet, alpha_generator_func=agen, set_alpha_scale=set_alpha_scale)
        x = sampler.sample(S=sc["S"], shape=(sc["batch"], 4, 64, 64), input=inputs, uc=uc, guidance_scale=sc["guidance"])
    finally:
        unet.restore_first_conv_from_SD = orig
        unet.undo_first_conv_restore()
        set_alpha_scale(unet, 1)
    _report(x, gold[name], f"sampler/{name}", 7.2e-3, 8.7e-3)  # <= 2x measured: 2.8e-3 (MIS S=10) / 3.7e-3 (PLMS S=4)
