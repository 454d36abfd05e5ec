#!/usr/bin/env python
"""bench.py -- images/sec/GPU of the InstanceDiffusion sampling hot path on B200.

Workload (BASELINE.json configs[1]): batch=4 images of 512x512 (latent 64x64), 8 box instances
each, 50-step PLMS, classifier-free guidance 7.5, alpha schedule [0.8, 0, 0.2], fp16 compute.
One "step" of the bench contract = one full `sampler.sample(...)` call over one batch (latent out);
timed region = the sampler only (no CLIP, no VAE), as SURVEY.md section 8d prescribes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mis 0.0] [--impl reference]

N > 1: launched under torch.distributed.run, one rank per GPU; rank 0's synthetic weights are
broadcast once over NCCL, every rank then samples its own batch of prompts (weak scaling, no
per-step collective).  `value` = images of all ranks / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from functools import partial

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "config2: batch=4 512x512, 8 box instances, 50-step PLMS, CFG 7.5, fp16, 1 GPU"
BATCH, N_INST, S_STEPS, GUIDANCE, ALPHA_TYPE = 4, 8, 50, 7.5, [0.8, 0.0, 0.2]
FLAVOR, LATENT = "box", 64
MIS_DEFAULT = 0.36  # inference.py:176

# BASELINE.json configs (per GPU; every config shards whole images over ranks, no per-step collective).
# Config 2 is the headline the metric is quoted on; the others are selectable parity / stress workloads.
CONFIGS = {
    2: dict(workload=WORKLOAD, batch=4, n=8, flavor="box", latent=64, mis=0.0),
    3: dict(workload="config3: batch=4 per GPU (32 over 8 GPUs) 512x512, box+point+scribble (test_scribble flags), "
                     "8 instances, 50-step PLMS, Multi-instance Sampler 0.36, alpha 0.8, fp16",
            batch=4, n=8, flavor="scribble", latent=64, mis=0.36),
    4: dict(workload="config4: batch=8 768x768 (latent 96x96), mask conditioning (test_mask flags), 16 instances, "
                     "50-step PLMS, bf16 storage (libidiff_b200_bf16.so), fp32 accumulation",
            batch=8, n=16, flavor="mask", latent=96, mis=0.0, dtype="bf16"),
    5: dict(workload="config5: batch=8 per GPU (64 over 8 GPUs) 512x512, 30 box instances, 50-step PLMS, "
                     "Multi-instance Sampler 0.36 (31 trajectories), fp16",
            batch=8, n=30, flavor="box", latent=64, mis=0.36),
}


def forwards_per_sample_call(S, n, mis):
    """UNet forwards per `sample()` call, each at batch B."""
    ms = int(S * mis)
    return 2 * ((n + 1) * (ms + 1) + (S - ms)) if mis > 0 else 2 * (S + 1)


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)
        return False

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the plain-torch restatement of the reference path on the host cores
# ------------------------------------------------------------------------------------------------
_CPU_STATE: dict = {}


def _best_thread_count() -> int:
    """All host threads the process may use -- unless oversubscription makes that slower (the GPU
    boxes expose 128 logical CPUs to a container with a smaller quota: 128 threads ran the same
    forward 20x slower than 8).  A 2-second matmul calibration picks the fastest count."""
    if "threads" in _CPU_STATE:
        return _CPU_STATE["threads"]
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, avail) if c <= avail})
    a = torch.randn(1536, 1536)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        a @ a
        t0 = time.perf_counter()
        for _ in range(3):
            a @ a
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
    _CPU_STATE["threads"] = best
    return best


def cpu_forward_seconds(n_forwards: int = 2, threads: int | None = None):
    """Times `n_forwards` steady-state UNet forwards (B=1, 512^2, fp32) of oracle/torch_oracle.py --
    the CPU restatement of the reference's forward_single_input -- after one untimed call."""
    from oracle import torch_oracle as TO
    from instancediffusion_b200 import synthetic
    from instancediffusion_b200.weights import UNIFUSION_FLAGS, synth_tensor, unet_config
    from instancediffusion_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
    threads = threads or _best_thread_count()
    torch.set_num_threads(threads)
    if "sd" not in _CPU_STATE:
        with torch.device("meta"):
            m = UNetModel(**unet_config("box"))
        _CPU_STATE["sd"] = {k: synth_tensor(k, tuple(v.shape), 0) for k, v in m.state_dict().items()
                            if "convnext" not in k}
    sd = _CPU_STATE["sd"]
    gb = synthetic.make_grounding_batch(1, N_INST, 3, "box")
    gi = dict(boxes=gb["boxes"], masks=gb["masks"], positive_embeddings=gb["text_embeddings"],
              scribbles=gb["scribbles"], polygons=gb["polygons"], segs=gb["segs"], points=gb["points"])
    x = synthetic.make_noise(1, 3)
    ctx = synthetic.make_context(1, 4)
    t = torch.full((1,), 601, dtype=torch.long)
    flags = UNIFUSION_FLAGS["box"]
    with torch.no_grad():
        TO.unet_forward(sd, x, t, ctx, gi, flags)
        t0 = time.perf_counter()
        for _ in range(n_forwards):
            TO.unet_forward(sd, x, t, ctx, gi, flags)
        dt = (time.perf_counter() - t0) / n_forwards
    return dt, threads


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write banners there (NCCL prints its version on
    communicator creation): point fd 1 at stderr for the duration of the run and keep the real stdout for the line."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict) -> None:
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path (the oracle port of its
    forward; the Python reference itself cannot travel to the GPU box), all host threads.  Each
    step is a bounded sample: `fw` steady-state forwards at B=1, extrapolated to the forward count
    of the workload (BASELINE.md section 4 'extrapolated')."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fpc = forwards_per_sample_call(S_STEPS, N_INST, args.mis)
    times = []
    threads = _best_thread_count()
    for i in range(args.warmup + args.steps):
        dt, threads = cpu_forward_seconds(1, threads)
        if i >= args.warmup:
            times.append(dt)
    t_fwd = sum(times) / len(times)
    # one sample() call of B images = fpc forwards at batch B; CPU time scales ~linearly in batch
    sec_per_image = fpc * t_fwd
    value = 1.0 / sec_per_image
    line = {
        "impl": "reference", "metric": "images/sec/GPU @512^2 fp16 50-step PLMS, 8 instances", "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": BATCH * sec_per_image * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "mis": args.mis, "forwards_per_call": fpc},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": f"{len(times)} steady-state UNet forwards at B=1 (fp32, {threads} threads), "
                                   f"{t_fwd:.2f} s each, x{fpc} forwards per image (extrapolated)"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)


# ------------------------------------------------------------------------------------------------
# the CUDA arm
# ------------------------------------------------------------------------------------------------
def build_pipeline(device, rank, world):
    from instancediffusion_b200 import parallel
    from instancediffusion_b200.ldm.models.diffusion.ldm import LatentDiffusion
    from instancediffusion_b200.weights import build_unet
    # rank 0 materialises the synthetic weights; the other ranks receive them over NCCL
    model = build_unet(FLAVOR, device, seed=0 if rank == 0 else None)
    torch.cuda.synchronize()
    parallel.barrier()
    t0 = time.perf_counter()
    from instancediffusion_b200 import ops
    sent = parallel.broadcast_module_(model, src=0, wire_dtype=ops.HALF)  # matrices in the 16-bit storage type (2.46 GB), vectors fp32
    torch.cuda.synchronize()
    parallel.barrier()
    bcast_ms = (time.perf_counter() - t0) * 1e3 if world > 1 else 0.0
    diffusion = LatentDiffusion(linear_start=0.00085, linear_end=0.012, timesteps=1000).to(device)
    # SD1.5 first conv swapped in at alpha == 0 (openaimodel.py:469-480).  The shipped 48 KB file is
    # a fixture under tests/golden/; a synthetic stand-in of the same shape is used if it is absent.
    p = os.path.join(ROOT, "tests", "golden", "sd15_first_conv.pt")
    if os.path.exists(p):
        sd_conv = torch.load(p, map_location="cpu")
    else:
        g = torch.Generator().manual_seed(5)
        sd_conv = {"weight": torch.randn((320, 4, 3, 3), generator=g) * 0.1, "bias": torch.zeros(320)}
    model.restore_first_conv_from_SD = lambda: (None if getattr(model, "_first_conv_restored", False)
                                                else model.set_sd_first_conv(sd_conv))
    return model, diffusion, sent, bcast_ms


def make_sampler(model, diffusion, mis):
    from instancediffusion_b200.ldm.models.diffusion.plms import PLMSSampler
    from instancediffusion_b200.ldm.models.diffusion.plms_instance import PLMSSamplerInst
    from instancediffusion_b200.utils.model import alpha_generator, set_alpha_scale
    agen = partial(alpha_generator, type=ALPHA_TYPE)
    if mis > 0:
        return PLMSSamplerInst(diffusion, model, alpha_generator_func=agen, set_alpha_scale=set_alpha_scale, mis=mis)
    return PLMSSampler(diffusion, model, alpha_generator_func=agen, set_alpha_scale=set_alpha_scale)


def host_inputs(model, seed, mis):
    """Pinned host copies of everything `sample()` consumes for one batch (the e2e leg copies them
    to the device inside the timed region)."""
    from instancediffusion_b200 import synthetic
    gti = model.grounding_tokenizer_input
    inputs, uc = synthetic.make_sampler_inputs(gti, BATCH, N_INST, seed, FLAVOR, mis=mis > 0, device="cpu", size=LATENT)
    lst = inputs if isinstance(inputs, list) else [inputs]

    def pin(t):
        if t.dim() == 4 and t.stride(-1) == 0:  # zero `segs` view: stays a broadcast view
            return t
        return t.contiguous().pin_memory()

    host = []
    for inp in lst:
        gi = {k: pin(v) for k, v in inp["grounding_input"].items()}
        host.append(dict(x=pin(inp["x"]), context=pin(inp["context"]), grounding_input=gi))
    return host, pin(uc), isinstance(inputs, list)


def to_device(host, uc, is_list, device, gti):
    nbytes = 0
    dev = []
    for h in host:
        gi = {}
        for k, v in h["grounding_input"].items():
            if v.dim() == 4 and v.stride(-1) == 0:
                gi[k] = torch.zeros((v.shape[0], v.shape[1], 1, 1), device=device).expand(*v.shape)
            else:
                gi[k] = v.to(device, non_blocking=True)
                nbytes += v.numel() * v.element_size()
        x = h["x"].to(device, non_blocking=True)
        c = h["context"].to(device, non_blocking=True)
        nbytes += x.numel() * 4 + c.numel() * 4
        gti.prepare({**gi, "text_embeddings": gi["positive_embeddings"]})
        dev.append(dict(x=x, timesteps=None, context=c, grounding_input=gi))
    ucd = uc.to(device, non_blocking=True)
    nbytes += uc.numel() * 4
    return (dev if is_list else dev[0]), ucd, nbytes


def roofline_pass(model, device, peaks):
    """One eager (graph-free) batched cond+uncond forward at the bench batch with every launch
    bracketed by CUDA events on the launching stream: per-kernel-class time, algorithmic FLOPs /
    bytes, and the roofline of the dominant class."""
    from instancediffusion_b200 import ops, synthetic
    from instancediffusion_b200.utils.model import set_alpha_scale
    gti = model.grounding_tokenizer_input
    inp, uc = synthetic.make_sampler_inputs(gti, BATCH, N_INST, 77, FLAVOR, mis=False, device=device, size=LATENT)
    inp["timesteps"] = torch.full((BATCH,), 601, dtype=torch.long, device=device)
    un = dict(x=inp["x"], timesteps=inp["timesteps"], context=uc)
    set_alpha_scale(model, 1)
    saved = model.use_cuda_graph
    model.use_cuda_graph = False
    model.forward_batched([inp, un])  # warm: hoisted tensors cached
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    agg = {}
    for _ in range(3):
        flush.zero_()  # > L2 (126 MB) written between iterations
        ops.PROFILE = []
        model.forward_batched([inp, un])
        torch.cuda.synchronize()
        for kind, fl, by, s, e in ops.PROFILE:
            a = agg.setdefault(kind, [0.0, 0.0, 0.0, 0])
            a[0] += s.elapsed_time(e) * 1e-3
            a[1] += fl
            a[2] += by
            a[3] += 1
        ops.PROFILE = None
    model.use_cuda_graph = saved
    tot = sum(a[0] for a in agg.values())
    breakdown = {k: {"share": a[0] / tot, "launches": a[3] // 3, "ms": a[0] / 3 * 1e3,
                     "tflops": (a[1] / a[0] / 1e12) if a[1] else None,
                     "gbs": a[2] / a[0] / 1e9} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    # kernel classes: every linear / conv3x3 / GEGLU launch is the same kernel template
    # (csrc/gemm2.cu gemm2_kernel<BN, MODE, TMA_EPI>), so they compete as one class for "dominant"
    fam = {"gemm2_kernel": [0.0, 0.0, 0.0, 0]}
    for k, a in agg.items():
        tgt = "gemm2_kernel" if k in ("gemm", "conv3x3", "gemm_geglu") else k
        f = fam.setdefault(tgt, [0.0, 0.0, 0.0, 0])
        for i in range(4):
            f[i] += a[i]
    traffic_tab = {}
    try:
        if (BATCH, LATENT) != (4, 64):  # the table was captured on forward batch 8 at 64x64 (configs 2 / 3)
            raise KeyError("no traffic capture for this workload")
        # DRAM bytes per launch per kernel family: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` over one
        # eager forward of this workload (tools/r2_profiles.sh -> tools/ncu_traffic.py)
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    except Exception:
        pass
    dom = max(fam.items(), key=lambda kv: kv[1][0])
    kind, a = dom
    traffic = traffic_tab.get(kind, {}).get("dram_bytes_per_launch")
    if a[1] > 0:
        achieved = a[1] / a[0] / 1e12
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        roof = {"bound": "tensor", "kernel": kind, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "launches_per_forward": a[3] // 3,
                "share_of_forward": a[0] / tot,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"}
    else:
        achieved = a[2] / a[0] / 1e9
        peak = peaks.get("hbm_gbs") or 6650.0
        roof = {"bound": "hbm", "kernel": kind, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"}
    # the north-star kernel (fused gated self-attention at the 64x64 level) reported next to it
    att = agg.get("attention_d40")
    if att and att[0] > 0:
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        roof["attention_d40"] = {"achieved": att[1] / att[0] / 1e12, "unit": "TFLOP/s", "frac": att[1] / att[0] / 1e12 / peak,
                                 "share_of_forward": att[0] / tot,
                                 "traffic": traffic_tab.get("attention2_kernel", {}).get("dram_bytes_per_launch")}
    n_launch = sum(a[3] for a in agg.values()) // 3
    return roof, breakdown, n_launch


def main():
    global WORKLOAD, BATCH, N_INST, FLAVOR, LATENT
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mis", type=float, default=None,
                    help="Multi-instance Sampler fraction of the headline leg (inference.py default 0.36); default: the "
                         "config's own (0 for config 2, whose plain-PLMS number is BASELINE's metric)")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[N-1]")
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dtype", default=None, choices=["fp16", "bf16"],
                    help="16-bit storage type (default: the config's own -- fp16, bf16 for config 4)")
    ap.add_argument("--no-mis-leg", action="store_true", help="skip the extra mis=0.36 leg of config 2")
    args = ap.parse_args()
    _claim_stdout()
    cfg = CONFIGS[args.config]
    WORKLOAD, BATCH, N_INST, FLAVOR, LATENT = cfg["workload"], cfg["batch"], cfg["n"], cfg["flavor"], cfg["latent"]
    if args.mis is None:
        args.mis = cfg["mis"]
    if args.impl == "reference":
        return run_reference_arm(args)

    from instancediffusion_b200 import _lib, ops, parallel
    from instancediffusion_b200.utils.model import set_alpha_scale
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CUDA arm has no CPU fallback (use --impl reference)")
    dtype_name = args.dtype or cfg.get("dtype", "fp16")
    ops.set_storage_dtype(torch.bfloat16 if dtype_name == "bf16" else torch.float16)
    _lib.load()
    rank, local_rank, world = parallel.init_distributed()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    model, diffusion, sent, bcast_ms = build_pipeline(device, rank, world)
    gti = model.grounding_tokenizer_input
    shape = (BATCH, 4, LATENT, LATENT)

    def reset():
        # every sample() call starts from a fresh model state and recomputes the per-sample hoisted
        # tensors (UniFusion tokens, object / text K/V): nothing is carried over between timed steps
        model.undo_first_conv_restore()
        set_alpha_scale(model, 1)
        model.clear_hoisted()

    def measure(mis, steps, warmup):
        """One leg: `warmup` untimed + `steps` timed sample() calls with inputs resident in HBM, then `steps`
        timed calls end to end (pinned host buffers in, latent back to the host inside the timed region).
        Device-timed with CUDA events, barrier + synchronize on both sides, max over ranks."""
        sampler = make_sampler(model, diffusion, mis)
        host, uc_host, is_list = host_inputs(model, 1000 + rank, mis)

        def run_resident(inputs, uc):
            # fresh trajectory state; the latent x is cloned so every step starts from the same noise
            if isinstance(inputs, list):
                ins = [dict(i, x=i["x"].clone()) for i in inputs]
            else:
                ins = dict(inputs, x=inputs["x"].clone())
            return sampler.sample(S=S_STEPS, shape=shape, input=ins, uc=uc, guidance_scale=GUIDANCE)

        dev_inputs, uc_dev, h2d_bytes = to_device(host, uc_host, is_list, device, gti)
        torch.cuda.synchronize()
        for _ in range(warmup):
            reset()
            run_resident(dev_inputs, uc_dev)
        torch.cuda.synchronize()
        parallel.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with ClockSampler(local_rank) as clk:
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(steps):
                reset()
                out = run_resident(dev_inputs, uc_dev)
            ev[1].record()
            torch.cuda.synchronize()
        parallel.barrier()
        t_dev = parallel.max_over_ranks(ev[0].elapsed_time(ev[1]) * 1e-3, device)
        assert torch.isfinite(out).all()
        result_host = torch.empty(shape, dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        parallel.barrier()
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev2[0].record()
        for _ in range(steps):
            reset()
            di, ud, _ = to_device(host, uc_host, is_list, device, gti)
            o = sampler.sample(S=S_STEPS, shape=shape, input=di, uc=ud, guidance_scale=GUIDANCE)
            result_host.copy_(o, non_blocking=True)
        ev2[1].record()
        torch.cuda.synchronize()
        parallel.barrier()
        t_e2e = parallel.max_over_ranks(ev2[0].elapsed_time(ev2[1]) * 1e-3, device)
        images = BATCH * steps * world
        return dict(value=images / t_dev, e2e=images / t_e2e, ms_per_step=t_dev / steps * 1e3, h2d=h2d_bytes,
                    d2h=result_host.numel() * 4, clocks=clk.summary(), fpc=forwards_per_sample_call(S_STEPS, N_INST, mis))

    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0
    head = measure(args.mis, args.steps, args.warmup)
    # The reference's stock sampler is the Multi-instance Sampler at mis=0.36 (inference.py:59-64,176): measured
    # in the same invocation (bounded: <= 3 steps) so that the driver sees both numbers.
    mis_leg = None
    if args.config == 2 and args.mis == 0 and not args.no_mis_leg:
        mis_leg = measure(MIS_DEFAULT, max(1, min(args.steps, 3)), 1)
    if rank != 0:
        return
    roof, breakdown, launches_per_fwd = roofline_pass(model, device, peaks)
    value, fpc = head["value"], head["fpc"]
    line = {
        "metric": "images/sec/GPU @512^2 fp16 50-step PLMS, 8 instances", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype_name, "data": "synthetic",
        "config": {"workload": WORKLOAD, "mis": args.mis, "global_batch": BATCH * world, "parallelism": f"dp{world}",
                   "forwards_per_call": fpc, "forward_batch": 2 * BATCH,
                   "l2_policy": "activations per forward (>1 GB at batch 8) exceed the 126 MB L2; roofline pass "
                                "flushes L2 (256 MB write) between iterations",
                   "cuda_graph": bool(model.use_cuda_graph), "weights": "seeded random (no checkpoint offline)",
                   "weight_broadcast_bytes": sent, "weight_broadcast_ms": bcast_ms,
                   "weight_broadcast_wire": f"{dtype_name} matrices + fp32 vectors, one NCCL broadcast at init"},
        "per_gpu_images_per_s": value / world,
        "clocks": head["clocks"],
        "e2e": {"value": head["e2e"], "unit": "images/s", "h2d_bytes_per_step": head["h2d"],
                "d2h_bytes_per_step": head["d2h"]},
        "gpu_launches": int(launches_per_fwd * (fpc // 2) * args.steps),
        "roofline": roof,
        "breakdown": breakdown,
    }
    # per-image algorithmic work: F_min(alpha=1)=1136, F_min(alpha=0)=803 GFLOP/forward/sample
    if args.config == 2 and args.mis == 0:
        tflop = (2 * 41 * 1.136 + 2 * 10 * 0.803)
        line["model_roofline"] = {"tflop_per_image_fmin": tflop, "achieved_tflops": value / world * tflop,
                                  "frac_of_sustained_peak": value / world * tflop / peak_tf}
    if mis_leg is not None:
        # mis=0.36, n=8: 2*[(n+1)*(ms+1) + (S-ms)] = 406 forwards per image; all MIS steps at alpha=1
        ms = int(S_STEPS * MIS_DEFAULT)
        n_a1 = (N_INST + 1) * (ms + 1) + (int(0.8 * S_STEPS) - ms)
        tflop = 2 * n_a1 * 1.136 + 2 * (S_STEPS - int(0.8 * S_STEPS)) * 0.803
        line["mis036"] = {"value": mis_leg["value"], "unit": "images/s", "per_gpu_images_per_s": mis_leg["value"] / world,
                          "e2e": mis_leg["e2e"], "ms_per_step": mis_leg["ms_per_step"], "forwards_per_call": mis_leg["fpc"],
                          "steps": max(1, min(args.steps, 3)), "warmup": 1, "clocks": mis_leg["clocks"],
                          "model_roofline": {"tflop_per_image_fmin": tflop,
                                             "achieved_tflops": mis_leg["value"] / world * tflop,
                                             "frac_of_sustained_peak": mis_leg["value"] / world * tflop / peak_tf}}
    if not args.no_cpu_baseline:
        try:
            dt, threads = cpu_forward_seconds(2)
            line["cpu_baseline"] = {
                "value": 1.0 / (fpc * dt), "unit": "images/s", "cores": threads, "kind": "port",
                "sample": f"2 steady-state UNet forwards at B=1 of oracle/torch_oracle.py (fp32, {threads} threads), "
                          f"{dt:.2f} s each, x{fpc} forwards per image (extrapolated)"}
        except Exception as exc:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc!r}"}
    _emit(line)


if __name__ == "__main__":
    try:
        main()
    finally:
        import torch.distributed as _dist
        if _dist.is_available() and _dist.is_initialized():
            _dist.destroy_process_group()
