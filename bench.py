#!/usr/bin/env python
"""bench.py — denoising-steps/s of the shape-branch denoiser (BASELINE.json metric, configs[1]).

One "step" = one classifier-free-guided DDIM step over 32 objects' 3x16^3 latents (of 64^3 SDFs): a UNet
evaluation at batch 64 ([uncond; cond]) + the fused CFG / x_prev update — the body of the reference's
DDIMSampler.ddim_sampling loop (samplers/ddim.py:154-177) at cfg2's batch.  Weights are random-init of the
reference architecture (413.5 M parameters), inputs synthetic (`data: synthetic`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); objects shard across ranks with no data-path collective
(SURVEY.md §8e), so scaling is weak: every rank denoises its own 32 objects and `value` counts all of them.

JSON line keys beyond the base contract: `roofline` (tensor-pipe fraction of the implicit-GEMM kernel, timed
live with CUDA events around every cs_conv3d launch of an instrumented step), `cpu_baseline` (the oracle — a
CPU restatement of the reference — on this host's cores, bounded sample), `e2e` (same step through the public
sampler API with HOST buffers in and out), `clocks`, `gpu_launches`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBJECTS = 32                      # cfg2: batch-32 sampling
UNET_GFLOP_PER_SAMPLE = 557.6     # SURVEY.md §8d (2*MAC, probe of the reference's own module)
METRIC = "denoising-steps/sec (64^3 SDF latent, bs32)"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = max((int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()), default=0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle (port of the reference) on host cores
# ------------------------------------------------------------------------------------------------
def cpu_oracle_steps_per_s(repeats: int, warmup: int):
    """Times the oracle's guided UNet evaluation for ONE object (batch 2 = [uncond; cond]) and scales linearly to the
    32-object step (the reference's CPU cost is linear in batch: BASELINE.md §4).  Returns (steps/s, cores, sample)."""
    import torch
    from oracle import denoiser as D, weights as Wt
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = D.UNET_FULL
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), 111)
    sched = D.register_schedule(**D.DIFFUSION)
    dd = D.ddim_schedule(sched, 100)
    g = torch.Generator().manual_seed(111)
    x = torch.randn(1, 3, 16, 16, 16, generator=g)
    c, uc = torch.randn(1, 1, 1280, generator=g), torch.randn(1, 1, 1280, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            D.p_sample_ddim(sd, cfg, dd, x, c, int(dd["timesteps"][-1]), 99, 3.0, uc)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per_object = sum(times) / len(times)
    sample = (f"oracle p_sample_ddim (fp32 torch CPU) on 1 object with CFG (UNet batch 2), mean of {repeats} after {warmup} warm-up, "
              f"x{OBJECTS} linear extrapolation to the {OBJECTS}-object step")
    return 1.0 / (per_object * OBJECTS), cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, sample = cpu_oracle_steps_per_s(repeats=max(1, args.steps), warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 guided DDIM step (UNet3DModel 413.5M params, 32 objects x CFG = batch 64, 3x16^3 latents of 64^3 SDFs), "
                               "reference algorithm on host CPU", "objects_per_step": OBJECTS},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from commonscenes_b200 import _lib, ops
    from commonscenes_b200.model.networks.diffusion_networks.network import DiffusionUNet
    from commonscenes_b200.model.networks.diffusion_networks.samplers.ddim import DDIMSampler
    from commonscenes_b200.model.sdfusion_txt2shape_model import UNET_PARAMS, diffusion_schedule

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_device()
    dev = torch.device("cuda", local)
    torch.manual_seed(111 + rank)

    with torch.device(dev):
        df = DiffusionUNet(dict(UNET_PARAMS), conditioning_key="crossattn")
        for p in df.parameters():                       # the reference zero-inits 18 convs: give them weights
            if p.dim() > 1 and float(p.abs().max()) == 0:
                torch.nn.init.normal_(p, std=0.02)
    df.eval()
    sched = diffusion_schedule()

    class Host:                                         # what DDIMSampler needs from SDFusionText2ShapeModel
        num_timesteps = 1000
        betas = sched["betas"].to(dev)
        alphas_cumprod = sched["alphas_cumprod"].to(dev)
    Host.df = df
    sampler = DDIMSampler(Host(), use_cuda_graph=True)
    sampler.make_schedule(100, ddim_eta=0.0, verbose=False)
    unet = df.diffusion_net
    steps_tab = sampler.ddim_timesteps[::-1]

    x = torch.randn(OBJECTS, 3, 16, 16, 16, device=dev)
    c = torch.randn(OBJECTS, 1, 1280, device=dev)
    uc = torch.randn(OBJECTS, 1, 1280, device=dev)
    ca = unet.context_vectors(torch.cat([uc, c]))
    t_dev = torch.empty(2 * OBJECTS, dtype=torch.int64, device=dev)

    def step(i, img):
        idx = 99 - (i % 100)
        t_dev.fill_(int(steps_tab[i % 100]))
        eps = sampler._eps(img, t_dev, ca)
        out, _ = ops.ddim_step(img, eps, guided=True, scale=3.0, a_t=float(sampler.ddim_alphas[idx]),
                               a_prev=float(sampler.ddim_alphas_prev[idx]), sigma=0.0,
                               sqrt_one_minus_at=float(sampler.ddim_sqrt_one_minus_alphas[idx]), want_pred_x0=False)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident loop (`value`) ----
    img = x
    for i in range(max(args.warmup, 3)):
        img = step(i, x)                                # restart from x_T each warm-up step: keeps values finite
    clocks = ClockSampler(local) if rank == 0 else None
    if clocks:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img = x
    for i in range(args.steps):
        img = step(i, img)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    kernels_per_step = sampler.kernels_per_eval + 1

    # ---- end to end through the public sampler call with host buffers (`e2e`) ----
    hx = torch.randn(OBJECTS, 3, 16, 16, 16).pin_memory()
    hc, huc = torch.randn(OBJECTS, 1, 1280).pin_memory(), torch.randn(OBJECTS, 1, 1280).pin_memory()
    hout = torch.empty(OBJECTS, 3, 16, 16, 16).pin_memory()
    h2d = hx.numel() * 4 + hc.numel() * 4 + huc.numel() * 4
    d2h = hout.numel() * 4

    ht = torch.empty(OBJECTS, dtype=torch.int64).pin_memory()
    h2d += ht.numel() * 8

    def e2e_step(i):
        # host buffers in -> the sampler's public single-step call (reference: DDIMSampler.p_sample_ddim) -> host buffer out
        ht.fill_(int(steps_tab[i % 100]))
        dx, dc, duc, dt = (h.to(dev, non_blocking=True) for h in (hx, hc, huc, ht))
        out, _ = sampler.p_sample_ddim(dx, dc, dt, index=99 - (i % 100), unconditional_guidance_scale=3.0,
                                       unconditional_conditioning=duc)
        hout.copy_(out, non_blocking=True)
    for i in range(3):
        e2e_step(i)
    barrier()
    k_e2e = max(3, min(args.steps, 20))
    e0.record()
    for i in range(k_e2e):
        e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    clock_rec = clocks.stop() if clocks else None

    # ---- roofline of the dominant kernel: CUDA events around every implicit-GEMM launch of one eager step ----
    prof = ops.ConvProfiler()
    with prof:
        unet(x, t_dev, context_vecs=ca, shared_prefix=True)      # the same evaluation the timed steps replay
    torch.cuda.synchronize()
    conv_ms, conv_tflop, n_conv = prof.summary()

    if world > 1:
        tt = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e = float(tt[0]), float(tt[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak_tf, peak_hbm, peak_src = _peaks()
    value = world * args.steps / (ms_total / 1e3)
    e2e_value = world * k_e2e / (ms_e2e / 1e3)
    achieved = conv_tflop / (conv_ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg2 guided DDIM step: UNet3DModel (413.5M params, random init) on 32 objects x CFG = batch 64 of 3x16^3 "
                               "latents of 64^3 SDFs + fused CFG/x_prev update; DDIM S=100 eta=0 scale=3",
                   "objects_per_gpu": OBJECTS, "global_objects": OBJECTS * world, "parallelism": f"objects sharded x{world}, no data-path collective",
                   "cache": "per-step working set (0.83 GB bf16 weights + activations) exceeds the 126 MB L2; no explicit flush",
                   "unet_tflop_per_step": 2 * OBJECTS * UNET_GFLOP_PER_SAMPLE / 1e3,
                   "conv_tflop_executed_per_step": conv_tflop,
                   "shared_prefix": "the layers in front of the first cross-attention see identical inputs in the uncond and cond "
                                    "halves of a guided step and are evaluated once (same values, same kernels): executed GEMM FLOPs "
                                    "are below the reference algorithm's 35.69 TFLOP",
                   "achieved_tflops_whole_step": 2 * OBJECTS * UNET_GFLOP_PER_SAMPLE / 1e3 / (ms_total / args.steps / 1e3),
                   "achieved_tflops_whole_step_note": "reference-algorithm FLOPs / time (throughput-equivalent, not executed FLOPs)"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": 448716288, "kernel": "cs::igemm_kernel (tcgen05 implicit GEMM)", "launches_per_step": n_conv,
                     "traffic_note": "dram__bytes_read+write of ONE representative launch from ncu --set full (profiles/r1e_igemm_shape3.txt: "
                                     "conv3d 448->448 @16^3 batch 64, 2841 GFLOP); its algorithmic bytes are 481 MB "
                                     "(235 MB in + 11 MB weights + 235 MB out), i.e. no re-reads from HBM",
                     "ms_per_step_in_kernel": conv_ms, "peak_source": f"bf16_tflops_sustained ({peak_src})",
                     "note": "achieved = algorithmic 2*MAC FLOPs of every cs_conv3d launch of one step / sum of their CUDA-event durations"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k_e2e},
        "gpu_launches": kernels_per_step * args.steps,
        "clocks": clock_rec,
    }
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_oracle_steps_per_s(repeats=3, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
