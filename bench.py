#!/usr/bin/env python
"""bench.py — denoising-steps/s of the shape-branch denoiser (BASELINE.json metric, configs[1]).

One "step" = one classifier-free-guided DDIM step over 32 objects' 3x16^3 latents (of 64^3 SDFs): a UNet
evaluation at batch 64 ([uncond; cond]) + the fused CFG / x_prev update — the body of the reference's
DDIMSampler.ddim_sampling loop (samplers/ddim.py:154-177) at cfg2's batch.  Weights are random-init of the
reference architecture (413.5 M parameters), inputs synthetic (`data: synthetic`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); objects shard across ranks with no data-path collective
(SURVEY.md §8e), so the headline scaling is weak: every rank denoises its own 32 objects and `value` counts all of them.

JSON line keys beyond the base contract: `roofline` (tensor-pipe fraction of the implicit-GEMM kernel, timed
live with CUDA events around every cs_conv3d launch of an instrumented step), `cpu_baseline` (the oracle — a
CPU restatement of the reference — on this host's cores, bounded sample), `e2e` (same step through the public
sampler API with HOST buffers in and out), `clocks`, `gpu_launches`, and two secondary blocks measured in the same
process after the headline loop:
  `train` — BASELINE cfg3/cfg4: the data-parallel denoiser training step (32 objects per rank; re-pack, forward,
            backward, bucketed NCCL all-reduce of the fp32 gradients overlapped with the backward, clip, AdamW) in ONE
            CUDA graph — the path that HAS a collective;
  `train_branch` — the same with everything around it: scene graphs -> GCN-E2 + rel_mlp -> frozen VQ-VAE encode of the 64^3
            SDFs -> denoiser -> gradient back into the graph networks -> both optimizers, also ONE CUDA graph;
  `cfg5`  — BASELINE cfg5's scene: 10 objects, guided DDIM S=100 + VQ-VAE decode to 64^3, the 20 forwards of each step
            split across the ranks (CFG-pair split, one all_gather of eps per step).

`--impl reference` (rank 0 only) times the reference's algorithm (the oracle port, pinned to the reference modules at max
|diff| 0) on the host cores: every timed step is ONE mini-batch of 7 objects with CFG (UNet batch 14) — the mini-batch
size the reference's own rel2shape uses (sdfusion_txt2shape_model.py:493-497) — and `value` scales it to the 32-object
step (`extrapolation`).  When a GPU is visible it also runs the same oracle ops eagerly on cuda:0 (`gpu_comparator`:
ATen/cuDNN kernels, TF32 and bf16 autocast, full batch 64, no extrapolation): what the reference's code reaches on this
GPU without this repo.
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
WORKLOAD = ("cfg2 guided DDIM step: UNet3DModel (413.5M params, random init) on 32 objects x CFG = batch 64 of 3x16^3 "
            "latents of 64^3 SDFs + fused CFG/x_prev update; DDIM S=100 eta=0 scale=3")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_igemm_dram_step_b64.json")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 6650.0, "fallback"


def _measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the implicit-GEMM launches of one step, from the committed ncu
    capture of tools/profile_step.py (summarised by tools/ncu_summary.py --dram).  None when the capture is absent."""
    try:
        with open(TRAFFIC_FILE) as f:
            t = json.load(f)
        return t
    except Exception:
        return None


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
REF_MINI_BATCH = 7               # the reference's rel2shape mini-batch (sdfusion_txt2shape_model.py:493)


def cpu_oracle_steps_per_s(repeats: int, warmup: int, objects: int = REF_MINI_BATCH):
    """Times the oracle's guided DDIM step (p_sample_ddim: UNet on [uncond; cond] + the x_prev update) for ONE mini-batch
    of `objects` objects and scales linearly to the 32-object step (the reference itself walks a 32-object step in
    mini-batches; its CPU cost is linear in batch: BASELINE.md §4).  Returns (steps/s, cores, sample text, seconds per
    timed mini-batch step)."""
    import torch
    from oracle import denoiser as D, weights as Wt
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = D.UNET_FULL
    sd = Wt.synth_state_dict(D.unet_param_shapes(cfg), 111)
    sched = D.register_schedule(**D.DIFFUSION)
    dd = D.ddim_schedule(sched, 100)
    g = torch.Generator().manual_seed(111)
    x = torch.randn(objects, 3, 16, 16, 16, generator=g)
    c, uc = torch.randn(objects, 1, 1280, generator=g), torch.randn(objects, 1, 1280, generator=g)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            D.p_sample_ddim(sd, cfg, dd, x, c, int(dd["timesteps"][-1]), 99, 3.0, uc)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per_mb = sum(times) / len(times)
    factor = OBJECTS / objects
    sample = (f"oracle p_sample_ddim (fp32 torch CPU, {cores} threads) on one mini-batch of {objects} objects with CFG (UNet batch "
              f"{2 * objects}), mean of {repeats} after {warmup} warm-up; x{factor:.3f} linear extrapolation to the {OBJECTS}-object step")
    return 1.0 / (per_mb * factor), cores, sample, per_mb


def gpu_comparator():
    """The reference's algorithm as plain PyTorch eager ops (ATen / cuDNN) on cuda:0, full batch 64: fp32 with TF32 (torch's
    default for convs, what the reference's own code would run) and bf16 autocast.  None without a GPU."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        from oracle import denoiser as D, weights as Wt
        cfg = D.UNET_FULL
        sd = {k: v.cuda() for k, v in Wt.synth_state_dict(D.unet_param_shapes(cfg), 111).items()}
        g = torch.Generator().manual_seed(0)
        B = 2 * OBJECTS
        x = torch.randn(B, 3, 16, 16, 16, generator=g).cuda()
        t = torch.full((B,), 500).cuda()
        ctx = torch.randn(B, 1, 1280, generator=g).cuda()
        torch.backends.cudnn.benchmark = True
        out = {"what": "oracle.denoiser.unet_forward (= the reference's op sequence) eager on cuda:0, UNet batch 64, "
                       "mean of 3 after 1 warm-up, CUDA events", "unit": "ms per guided UNet evaluation"}
        with torch.device("cuda"), torch.no_grad():
            for name, ac in (("fp32_tf32", False), ("bf16_autocast", True)):
                torch.backends.cuda.matmul.allow_tf32 = True
                torch.backends.cudnn.allow_tf32 = True
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                    D.unet_forward(sd, cfg, x, t, ctx)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(3):
                        D.unet_forward(sd, cfg, x, t, ctx)
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                out[name] = {"ms": ms, "steps_per_s": 1000.0 / ms}
        return out
    except Exception as e:          # the CPU number must not depend on the comparator
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    objs = int(os.environ.get("CS_REF_OBJECTS", REF_MINI_BATCH))
    val, cores, sample, per_mb = cpu_oracle_steps_per_s(repeats=max(1, args.steps), warmup=max(1, args.warmup), objects=objs)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * per_mb, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "objects_per_gpu": OBJECTS, "global_objects": OBJECTS,
                   "arm": "reference algorithm (oracle port) on the host CPU"},
        "extrapolation": {"objects_timed_per_step": objs, "unet_batch_timed": 2 * objs, "factor": OBJECTS / objs,
                          "ms_per_full_step": 1000.0 / val,
                          "note": "ms_per_step is the MEASURED time of one timed step (one reference-sized mini-batch); value = "
                                  "1 / (ms_per_step * factor): steps/s of the 32-object step"},
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_gpu_comparator:
        line["gpu_comparator"] = gpu_comparator()
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from commonscenes_b200 import _lib, ops, parallel
    from commonscenes_b200.model.sdfusion_txt2shape_model import default_opt
    from commonscenes_b200.model.VAEGAN_V2FULL import Sg2ScVAEModel
    from commonscenes_b200.train import ShapeBranchTrainStep

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _lib.require_device()
    dev = torch.device("cuda", local)
    torch.manual_seed(111)                              # identical weights on every rank (data-parallel replicas)

    # the v2_full scene model around the denoiser (GCN-E2 + rel_mlp + embeddings: 11 M parameters) -- its shape branch is what
    # the `train_branch` block steps; every other block uses its .Diff (denoiser + frozen VQ-VAE, reference wiring)
    vocab = {"object_idx_to_name": [f"o{i}" for i in range(36)], "pred_idx_to_name": [f"p{i}" for i in range(16)]}
    scene = Sg2ScVAEModel(vocab, diff_opt=default_opt(device=f"cuda:{local}"), embedding_dim=64, mlp_normalization="batch",
                          residual=True, gconv_num_layers=5).to(torch.device("cuda", local))
    model = scene.Diff
    df = model.df
    with torch.no_grad():
        for p in df.parameters():                       # the reference zero-inits 18 convs: give them weights
            if p.dim() > 1 and float(p.abs().max()) == 0:
                torch.nn.init.normal_(p, std=0.02)
    df.eval()
    torch.manual_seed(111 + rank)                       # per-rank inputs
    sampler = model.ddim_sampler
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

    def max_over_ranks(*vals):
        if world == 1:
            return vals
        tt = torch.tensor(vals, device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tuple(float(v) for v in tt)

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
    ops.conv3d_variant_counts(reset=True)
    with prof:
        unet(x, t_dev, context_vecs=ca, shared_prefix=True)      # the same evaluation the timed steps replay
    torch.cuda.synchronize()
    conv_ms, conv_tflop, n_conv = prof.summary()
    variants = ops.conv3d_variant_counts(reset=True)

    # ---- cfg5: one 10-object scene, guided DDIM S=100 + decode, the 2 x 10 forwards of a step split over the ranks ----
    cfg5 = None
    if not args.no_cfg5:
        n5 = 10
        g5 = torch.Generator(device=dev).manual_seed(5)
        data5 = {"sdf": torch.zeros(n5, 1, 64, 64, 64, device=dev), "rel": torch.randn(n5, 1, 1280, device=dev, generator=g5),
                 "uc": torch.randn(n5, 1, 1280, device=dev, generator=g5)}
        parallel.rel2shape_pair_sharded(model, data5, ddim_steps=5, uc_scale=3.0, seed=7)          # warm-up: graph capture, VQ-VAE packs
        barrier()
        e0.record()
        sdf5 = parallel.rel2shape_pair_sharded(model, data5, ddim_steps=100, uc_scale=3.0, seed=7)
        e1.record()
        barrier()
        (ms5,) = max_over_ranks(e0.elapsed_time(e1))
        # the same scene with the full 1000-step ancestral DDPM chain BASELINE cfg5 names (not a reference capability: the
        # reference samples with DDIM only; parity of this sampler is unpinned by construction)
        parallel.rel2shape_pair_sharded(model, data5, uc_scale=3.0, seed=7, sampler="ddpm", ddpm_timesteps=10)     # warm-up
        barrier()
        e0.record()
        sdf5p = parallel.rel2shape_pair_sharded(model, data5, uc_scale=3.0, seed=7, sampler="ddpm")
        e1.record()
        barrier()
        (ms5p,) = max_over_ranks(e0.elapsed_time(e1))
        units = [hi - lo for lo, hi in parallel.pair_units(n5, world)]
        cfg5 = {"workload": "BASELINE cfg5 scene: 10 objects, guided DDIM S=100 (scale 3) + VQ-VAE decode to 64^3; ddpm_*: the same scene "
                            "with 1000 ancestral DDPM steps", "seconds": ms5 / 1e3,
                "objects_per_s": n5 / (ms5 / 1e3), "ddpm_seconds": ms5p / 1e3, "ddpm_object_steps_per_s": 1000 * n5 / (ms5p / 1e3),
                "ddpm_finite": bool(torch.isfinite(sdf5p).all()), "forwards_per_rank_per_step": units,
                "collective": "one all_gather of eps (48 KiB per forward) per step" if world > 1 else "none (one rank)",
                "finite": bool(torch.isfinite(sdf5).all()), "shape": list(sdf5.shape)}
        del sdf5, sdf5p, data5

    # ---- train: the data-parallel denoiser training step (cfg3 / cfg4), one CUDA graph incl. the NCCL all-reduces ----
    train = train_branch = None
    if not args.no_train:
        per_rank = 32
        branch = ShapeBranchTrainStep(scene)
        stepper = branch.denoiser
        z = torch.randn(per_rank, 3, 16, 16, 16, device=dev)
        ctx = torch.randn(per_rank, 1, 1280, device=dev)
        stepper.capture(per_rank, 1280)
        for _ in range(3):
            stepper.step_graphed(z, ctx)
        barrier()
        k_train = 10
        e0.record()
        for _ in range(k_train):
            loss_t, _ = stepper.step_graphed(z, ctx)
        e1.record()
        barrier()
        (ms_train,) = max_over_ranks(e0.elapsed_time(e1) / k_train)
        chk = stepper.flat_p.double().sum().reshape(1)
        same = True
        if world > 1:
            lo_, hi_ = chk.clone(), chk.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            same = bool(torch.equal(lo_, hi_))
        train = {"workload": "BASELINE cfg3/cfg4 denoiser training step: 32 objects per rank, weight re-pack + forward + backward + "
                             "bucketed gradient all-reduce + clip 5.0 + AdamW in one CUDA graph",
                 "ms_per_step": ms_train, "steps_per_s": 1000.0 / ms_train, "objects_per_s": per_rank * world * 1000.0 / ms_train,
                 "objects_per_rank": per_rank, "allreduce_bytes": int(stepper.flat_g.numel()) * 4 if world > 1 else 0,
                 "allreduce_dtype": "f32", "bucket_mb": 256, "buckets": len(stepper.buckets),
                 "algorithmic_tflops": 3 * per_rank * world * UNET_GFLOP_PER_SAMPLE / ms_train,
                 "replicas_identical": same, "loss": float(loss_t), "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
        del stepper.graph
        stepper.graph = None

        # ---- train_branch: the WHOLE v2_full shape-branch iteration (cfg3 per rank, cfg4 at 8 ranks) as one CUDA graph ----
        # per rank 4 scenes x 8 objects with 12 edges each; every rank runs the (tiny) graph networks on the global batch so that
        # BatchNorm statistics equal the single-process ones, and the denoiser on its own 32 objects
        scenes, per_scene, edges = 4 * world, 8, 12
        gb = torch.Generator().manual_seed(4242)                      # identical batch on every rank
        O_g, T_g = scenes * per_scene, scenes * edges
        sub = torch.randint(0, per_scene, (scenes, edges), generator=gb)
        obj = (sub + torch.randint(1, per_scene, (scenes, edges), generator=gb)) % per_scene
        off = (torch.arange(scenes) * per_scene).view(-1, 1)
        triples = torch.stack([sub + off, torch.randint(1, 16, (scenes, edges), generator=gb), obj + off], dim=-1).view(T_g, 3).to(dev)
        objs = torch.randint(1, 36, (O_g,), generator=gb).to(dev)
        zl, text, relf = (torch.randn(n, d_, generator=gb).to(dev) for n, d_ in ((O_g, 64), (O_g, 512), (T_g, 512)))
        sdfs = (torch.randn(O_g, 1, 64, 64, 64, generator=gb) * 0.1).clamp_(-0.2, 0.2).to(dev)
        branch.capture(O_g, T_g)
        for _ in range(3):
            branch.step_graphed(zl, objs, triples, text, relf, sdfs)
        barrier()
        e0.record()
        for _ in range(k_train):
            loss_b, _ = branch.step_graphed(zl, objs, triples, text, relf, sdfs)
        e1.record()
        barrier()
        (ms_branch,) = max_over_ranks(e0.elapsed_time(e1) / k_train)
        train_branch = {"workload": "BASELINE cfg3 (per rank) / cfg4 (8 ranks): whole v2_full shape-branch iteration -- scene graphs (4 scenes x "
                                    "8 objects, 12 edges each, per rank) -> GCN-E2 + rel_mlp (train-mode BatchNorm) -> frozen VQ-VAE encode of 64^3 "
                                    "SDFs -> denoiser fwd/bwd -> gradient back into the graph networks -> clip + AdamW on both groups, one CUDA graph",
                        "ms_per_step": ms_branch, "objects_per_s": per_rank * world * 1000.0 / ms_branch, "scenes_per_s": scenes * 1000.0 / ms_branch,
                        "global_objects": O_g, "global_triples": T_g, "objects_per_rank": per_rank,
                        "collectives": (f"{len(stepper.buckets)} bucketed all-reduces of the denoiser gradients ({int(stepper.flat_g.numel()) * 4} B fp32) "
                                        f"+ 1 of the graph-side gradients ({int(branch.graph_params.flat_g.numel()) * 4} B)")
                                       if world > 1 else "none (one rank)",
                        "loss": float(loss_b), "finite": bool(torch.isfinite(loss_b))}

    ms_total, ms_e2e = max_over_ranks(ms_total, ms_e2e)

    def leave():
        """Multi-rank exit.  The data-parallel training step lives in CUDA graphs that hold NCCL kernels; tearing the
        communicator down under them (destroy_process_group) blocked forever on a 2-GPU box (gpurun_out/r2e_bench_2gpu.log:
        the JSON line was printed, then the job sat until the time limit).  Every rank therefore meets at a last barrier,
        drains its device and leaves without running the NCCL teardown."""
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        leave()
        return

    peak_tf, peak_hbm, peak_src = _peaks()
    value = world * args.steps / (ms_total / 1e3)
    e2e_value = world * k_e2e / (ms_e2e / 1e3)
    achieved = conv_tflop / (conv_ms / 1e3)
    traffic = _measured_traffic()
    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "objects_per_gpu": OBJECTS, "global_objects": OBJECTS * world, "parallelism": f"objects sharded x{world}, no data-path collective",
                   "cache": "per-step working set (0.83 GB bf16 weights + activations) exceeds the 126 MB L2; no explicit flush",
                   "unet_tflop_per_step": 2 * OBJECTS * UNET_GFLOP_PER_SAMPLE / 1e3,
                   "conv_tflop_executed_per_step": conv_tflop,
                   "exact_identities": "executed GEMM FLOPs are below the reference algorithm's 35.69 TFLOP: (1) the layers in front of "
                                       "the first cross-attention see identical inputs in the uncond and cond halves of a guided step "
                                       "and are evaluated once; (2) nearest-upsample + 3x3x3 conv runs as four 3x2x2 phase convs on the "
                                       "low-resolution tensor (12 of 27 taps) when that path is enabled",
                   "achieved_tflops_whole_step": 2 * OBJECTS * UNET_GFLOP_PER_SAMPLE / 1e3 / (ms_total / args.steps / 1e3),
                   "achieved_tflops_whole_step_note": "reference-algorithm FLOPs / time (throughput-equivalent, not executed FLOPs)",
                   "conv_kernel_variants": variants},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": None if traffic is None else traffic.get("dram_bytes_per_step"),
                     "kernel": "cs::igemm_kernel / cs::igemm2_kernel (tcgen05 implicit GEMM)", "launches_per_step": n_conv,
                     "traffic_note": ("no committed ncu capture found" if traffic is None else
                                      f"dram__bytes_read.sum + dram__bytes_write.sum summed over the {traffic.get('launches')} implicit-GEMM "
                                      f"launches of ONE guided step (ncu, {os.path.relpath(TRAFFIC_FILE, ROOT)}); algorithmic bytes of the "
                                      "same launches (inputs + weights + outputs, bf16): "
                                      f"{traffic.get('algorithmic_bytes_per_step')}"),
                     "ms_per_step_in_kernel": conv_ms, "peak_source": f"bf16_tflops_sustained ({peak_src})",
                     "note": "achieved = executed 2*MAC FLOPs of every cs_conv3d launch of one step / sum of their CUDA-event durations"},
        "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": k_e2e},
        "gpu_launches": kernels_per_step * args.steps,
        "clocks": clock_rec,
    }
    if train is not None:
        line["train"] = train
    if train_branch is not None:
        line["train_branch"] = train_branch
    if cfg5 is not None:
        line["cfg5"] = cfg5
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_oracle_steps_per_s(repeats=3, warmup=1)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    leave()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-comparator", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
