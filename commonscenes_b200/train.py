"""Native training step of the shape-branch denoiser (SURVEY.md §8 a10-a11, §8e "training partition").

What the reference does per iteration (train_3dfront.py:345-418 with VAEGAN_V2FULL.py:642-650):
    z = VQVAE.encode(sdf) (frozen) ; t ~ U[0, T) ; x_t = q_sample(z, t, eps) ; loss = 100 * mse(UNet(x_t, t, c), eps)
    loss.backward() ; clip_grad_norm_(df_module.parameters(), 5.0) ; AdamW.step()
and, under DistributedDataParallel, an all-reduce (mean) of the gradients across ranks.

`DenoiserTrainStep` is that iteration on the B200 kernels without autograd or DDP wrappers:
  * all trainable parameters of the denoiser are re-homed into ONE flat fp32 buffer (the nn.Parameters become views, so
    state_dict()/load_state_dict() and the reference's checkpoint format are unchanged), with matching flat buffers for
    the gradient and the AdamW moments;
  * forward_train / backward are the explicit kernels of unet_train.py, writing gradients straight into the flat buffer;
  * multi-GPU: one NCCL all-reduce per gradient bucket on a side stream, issued as soon as the backward has passed the
    bucket's blocks (buckets follow the UNet's block order, i.e. the reverse of the backward), then ONE cs_sumsq + ONE
    cs_adamw launch over the flat buffers (clip factor computed on the device, no host sync).
The conditioning gradient d_c is returned so the caller can continue into rel_mlp / GCN-E2.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import ctypes as C
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib, ops, ops_bwd
from .model.networks.diffusion_networks.unet_train import GradSink


def make_buckets(sizes: List[int], limit: int) -> List[tuple]:
    """Contiguous [lo, hi) element ranges of the flat gradient buffer, each closed as soon as it holds >= limit elements
    (parameter order; the backward finishes them from the last to the first)."""
    buckets, start, acc = [], 0, 0
    for n in sizes:
        acc += n
        if acc >= limit:
            buckets.append((start, start + acc))
            start, acc = start + acc, 0
    if acc:
        buckets.append((start, start + acc))
    return buckets


def fused_gradient_layout(sizes: List[int], slots: List[Optional[int]]):
    """Flat-buffer layout of the fused-update mode, from the per-parameter sizes (elements, module order) and, for the
    packed-gradient weights, their slot sizes (None = plain parameter).  Parameters sit [plain ... | packed ...] (each part
    in module order) in flat_p / flat_m / flat_v and, with the slot sizes, in flat_g.  Returns (p_off, g_off, final_from,
    n_plain, order): `final_from[i]` = the flat_g offset from which everything is final once the backward -- which walks the
    parameters from the last to the first -- has passed parameter i: the start of the first packed slot at or after i
    (the plain region in front of the slots is only final when the whole backward is)."""
    n = len(sizes)
    order = [i for i in range(n) if slots[i] is None] + [i for i in range(n) if slots[i] is not None]
    p_off, g_off = [0] * n, [0] * n
    po = go = 0
    for i in order:
        p_off[i], g_off[i] = po, go
        po += sizes[i]
        go += sizes[i] if slots[i] is None else slots[i]
    n_plain = sum(sizes[i] for i in range(n) if slots[i] is None)
    final_from, nxt = [0] * n, go
    for i in reversed(range(n)):
        if slots[i] is not None:
            nxt = g_off[i]
        final_from[i] = nxt
    return p_off, g_off, final_from, n_plain, order


def flat_optimizer_state_dict(params, offsets, flat_m, flat_v, step: int, hp: dict) -> dict:
    """The flat AdamW moments in torch.optim.AdamW.state_dict() layout -- what the reference checkpoints hold
    (SDFusionText2ShapeModel.save(save_opt=True): sdfusion_txt2shape_model.py:636-650; VAE.save: VAE.py:334-340) -- so a run
    can move between torch's optimizer and the native step in either direction."""
    state = {}
    for i, p in enumerate(params):
        off, n = offsets[p], p.numel()
        state[i] = {"step": torch.tensor(float(step)), "exp_avg": flat_m[off:off + n].view(p.shape).clone(),
                    "exp_avg_sq": flat_v[off:off + n].view(p.shape).clone()}
    group = dict(lr=hp["lr"], betas=tuple(hp["betas"]), eps=hp["eps"], weight_decay=hp["weight_decay"], amsgrad=False,
                 maximize=False, foreach=None, capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True,
                 params=list(range(len(params))))
    return {"state": state, "param_groups": [group]}


def load_flat_optimizer_state(sd: dict, params, offsets, flat_m, flat_v) -> int:
    """Inverse of flat_optimizer_state_dict for a torch.optim.AdamW state dict over the SAME parameter order; returns the
    step count (all parameters of a group share it).  Parameters without state (never stepped) keep zero moments."""
    ids = [i for g in sd["param_groups"] for i in g["params"]]
    if len(ids) != len(params):
        raise ValueError(f"optimizer state covers {len(ids)} parameters, the model has {len(params)}")
    step = 0
    flat_m.zero_(); flat_v.zero_()
    for i, p in zip(ids, params):
        st = sd["state"].get(i)
        if st is None:
            continue
        off, n = offsets[p], p.numel()
        if tuple(st["exp_avg"].shape) != tuple(p.shape):
            raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(p.shape)}")
        flat_m[off:off + n].view(p.shape).copy_(st["exp_avg"])
        flat_v[off:off + n].view(p.shape).copy_(st["exp_avg_sq"])
        step = max(step, int(float(st["step"])))
    return step


class _RepackEntry(C.Structure):
    """cs_repack_entry of include/cs_b200.h."""
    _fields_ = [("p_off", C.c_int64), ("g_off", C.c_int64), ("first_tile", C.c_int64), ("fwd", C.c_void_p), ("dgrad", C.c_void_p),
                ("Cout", C.c_int32), ("Cin", C.c_int32), ("taps", C.c_int32), ("C1", C.c_int32), ("group", C.c_int32), ("reserved", C.c_int32)]


class DenoiserTrainStep:
    def __init__(self, diff_model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 5.0, loss_scale: float = 100.0, group: Optional[dist.ProcessGroup] = None,
                 bucket_mb: int = 256, fused_update: Optional[bool] = None):
        """diff_model: SDFusionText2ShapeModel mirror (uses .df.diffusion_net, the schedule tables and q_sample).
        Defaults follow the reference: AdamW lr 1e-4 (VAEGAN_V2FULL.py:642-650; torch's default betas/eps/decay), clip 5.0
        (train_3dfront.py:399), total loss weight 100 on loss_df (train_3dfront.py:387).

        fused_update (default: on for CUDA parameters): the GEMM-class weights that have both kernel-layout packs keep their
        gradient in the weight-gradient kernel's packed layout and are updated by ONE cs_adamw_repack launch (clip + AdamW +
        both bf16 re-packs + clearing the gradient), instead of un-pack / AdamW / two pack kernels per weight.  The flat
        buffers then hold the remaining ("plain") parameters first and those weights after them; `offsets` is still the
        position of every parameter in flat_p / flat_m / flat_v, `flat_g` is laid out [plain | packed slots]."""
        self.model = diff_model
        self.unet = diff_model.df.diffusion_net
        self.trainer = self.unet.trainer()
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.max_grad_norm, self.loss_scale = max_grad_norm, loss_scale
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0
        params = [p for p in self.unet.parameters() if p.requires_grad]
        dev = params[0].device
        self.fused = (dev.type == "cuda") if fused_update is None else (bool(fused_update) and dev.type == "cuda")
        regular = self._discover_packed_weights(params) if self.fused else {}
        self.fused = bool(regular)
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # keep every view 16-byte aligned
        slots = [regular[p]["slot"] if p in regular else None for p in params]
        p_off, g_off, final_from, self.n_plain, order = fused_gradient_layout(sizes, slots)
        packed = [p for p in params if p in regular]
        total = sum(sizes)
        g_sizes = [sizes[i] if slots[i] is None else slots[i] for i in range(len(params))]
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(sum(g_sizes), dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: Dict[nn.Parameter, torch.Tensor] = {}
        self.packed_views: Dict[nn.Parameter, torch.Tensor] = {}
        self.offsets: Dict[nn.Parameter, int] = {}
        self.g_offsets: Dict[nn.Parameter, int] = {}
        self._final_from: Dict[nn.Parameter, int] = {}
        with torch.no_grad():
            for i in order:
                p, off, goff = params[i], p_off[i], g_off[i]
                v = self.flat_p[off:off + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                if p in regular:
                    r = regular[p]
                    self.packed_views[p] = self.flat_g[goff:goff + r["slot"]].view(r["Cout"], r["taps"], r["ctot"])
                else:
                    self.views[p] = self.flat_g[goff:goff + p.numel()].view(p.shape)
                self.offsets[p], self.g_offsets[p] = off, goff
                # without packed slots flat_g is laid out like flat_p and a parameter's own offset is the bound
                self._final_from[p] = final_from[i] if self.fused else goff
        self.params = params
        self.step_dev = torch.zeros((), dtype=torch.int32, device=dev)     # device copy of step_count (graph replays)
        self.graph = None
        self.sumsq = torch.zeros((), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        # gradient buckets in flat_g order; the backward fills them from the last to the first
        self.buckets: List[tuple] = make_buckets([g_sizes[i] for i in order], bucket_mb * (1 << 20) // 4)
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.unet._packed = None        # parameters moved: rebuild the kernel-layout copies
        if self.fused:
            self._build_repack_table(packed, regular)
            self.refresh_packs()

    # ------------------------------------------------------------------------------------------
    # fused optimizer: packed-gradient weights
    # ------------------------------------------------------------------------------------------
    def _discover_packed_weights(self, params) -> dict:
        """The parameters whose forward AND data-gradient kernel layouts are plain re-orderings of the parameter itself
        (ResBlock / Down / Upsample convs, skip 1x1x1s, proj_in / proj_out, to_out, ff.net[2]): they own one cached "fwd" and
        one "dgrad" pack buffer after a pack pass.  Weights packed through a re-arranged temporary (q|k|v, GEGLU rows,
        up-sample phase filters, the few-channel stem / head) stay on the un-pack + cs_adamw + cs_pack_weight path."""
        import ast
        from . import ops as _ops
        self.unet._packed = None
        self.trainer._ensure()
        by_id = {id(p): p for p in params}
        found: Dict[nn.Parameter, dict] = {}
        for (kind, pid, shape), (ref, buf) in list(_ops._PACK_BUFFERS.items()):
            p = by_id.get(pid)
            if p is None or ref() is not p or buf.device != p.device:
                continue
            d = found.setdefault(p, {"fwd": []})
            if kind == "dgrad":
                d["dgrad"] = (buf, (kind, pid, shape))
            elif kind.startswith("fwd"):
                d["fwd"].append((buf, (kind, pid, shape), ast.literal_eval(kind[3:])))
        regular = {}
        for p, d in found.items():
            if "dgrad" not in d or len(d["fwd"]) != 1:
                continue
            fwd, fkey, parts = d["fwd"][0]
            co, ci = p.shape[0], p.shape[1]
            taps = p.numel() // (co * ci)
            ctot = fwd.shape[2]
            if co % 8 or ci % 8 or parts[0] % 8 or tuple(fwd.shape) != (co, taps, ctot) or tuple(d["dgrad"][0].shape) != (ci, taps, _ops._pad64(co)):
                continue
            regular[p] = dict(fwd=fwd, dgrad=d["dgrad"][0], keys=(fkey, d["dgrad"][1]), Cout=co, Cin=ci, taps=taps, C1=int(parts[0]),
                              ctot=ctot, slot=co * taps * ctot)
        return regular

    def _build_repack_table(self, packed, regular) -> None:
        entries = (_RepackEntry * len(packed))()
        self._repack_tile_ci = 16
        tile = 0
        self._regular_keys = []
        for i, p in enumerate(packed):
            r = regular[p]
            gmax = max(1, 27 // r["taps"])       # a tile = 16 co x (16 * group) ci x taps: <= 432 cells per output channel;
            c16 = (r["Cin"] + 15) // 16          # equal-width tiles (448 input channels of a 1x1x1 weight -> 2 x 224, not 432 + 16)
            group = -(-c16 // -(-c16 // gmax))
            entries[i] = _RepackEntry(self.offsets[p], self.g_offsets[p], tile, r["fwd"].data_ptr(), r["dgrad"].data_ptr(),
                                      r["Cout"], r["Cin"], r["taps"], r["C1"], group, 0)
            tile += ((r["Cout"] + 15) // 16) * ((r["Cin"] + 16 * group - 1) // (16 * group))
            self._regular_keys += [(k, p) for k in r["keys"]]
        raw = torch.frombuffer(bytearray(bytes(entries)), dtype=torch.uint8).clone()
        self._repack_table = raw.to(self.flat_p.device)
        self._repack_n, self._repack_tiles = len(packed), tile
        self._repack_taps = max(regular[p]["taps"] for p in packed)
        self._repack_params = packed
        self._repack_keep = [regular[p]["fwd"] for p in packed] + [regular[p]["dgrad"] for p in packed]     # keep the buffers alive
        self._regular_versions = None

    def refresh_packs(self) -> None:
        """Rebuild the bf16 packs of the fused-update weights from the fp32 master weights with the stand-alone pack kernels
        (after anything other than cs_adamw_repack changed them: construction, load_state_dict, a rolled-back warm-up)."""
        if not self.fused:
            self.unet._packed = None
            return
        from . import ops as _ops
        for key, _ in self._regular_keys:
            _ops._PACK_MAINTAINED.pop(key, None)
        self.unet._packed = None
        self.trainer._ensure()
        for key, p in self._regular_keys:
            _ops._pack_maintain(key, p)
        self._regular_versions = sum(p._version for p in self._repack_params)

    def __del__(self):
        try:        # the packs of these weights are no longer kept current by anybody
            from . import ops as _ops
            for key, _ in getattr(self, "_regular_keys", []):
                _ops._PACK_MAINTAINED.pop(key, None)
        except Exception:
            pass

    def _check_packs_current(self) -> None:
        if self.fused and sum(p._version for p in self._repack_params) != self._regular_versions:
            self.refresh_packs()        # somebody wrote the parameters through torch (load_state_dict, an initialiser)

    # ------------------------------------------------------------------------------------------
    def _allreduce_ready(self, done_off: int, pending: List[int]):
        """Launch the all-reduce of every bucket that lies entirely at or after `done_off` (already final)."""
        while pending and self.buckets[pending[-1]][0] >= done_off:
            lo, hi = self.buckets[pending.pop()]
            if self.comm_stream is None:        # host tensors (the gloo CPU tests of this scheduling logic)
                dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                continue
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)

    def step(self, z: torch.Tensor, cond: torch.Tensor, t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
             need_dcond: bool = False, grad_weight: float = 1.0):
        """One optimisation step on this rank's shard.  z: (B, 3, 16, 16, 16) fp32 latents (VQVAE encode_only output),
        cond: (B, 1, context_dim) fp32.  Returns (loss tensor on the device [mean MSE, unscaled], d_cond or None).
        grad_weight: this rank's share of the global batch relative to an equal split (n_local * world / n_total), so that
        the 1/world mean over ranks is the global-batch mean even when the shards are ragged.  B == 0 (more ranks than
        objects) contributes zero gradients but still takes part in every collective."""
        m = self.model
        B = z.shape[0]
        dev = z.device
        if B == 0:
            return self._step_without_objects(cond)
        if t is None:
            t = torch.randint(0, m.num_timesteps, (B,), device=dev).long()
        if noise is None:
            noise = torch.randn_like(z)
        self._check_packs_current()
        x_t = m.q_sample(z, t, noise)
        eps, tape = self.trainer.forward_train(x_t, t, cond)
        self.loss.zero_()
        d_eps = ops_bwd.mse_loss_grad(eps, noise.float().contiguous(), self.loss, loss_scale=self.loss_scale * grad_weight)
        self._zero_grads()
        sink = GradSink(self.views, self.packed_views)
        if self.world > 1:
            pending = list(range(len(self.buckets)))
            # the backward visits layers in reverse parameter order: once a layer is done, its gradients and those of every
            # later parameter are final, so the buckets lying wholly beyond that offset can be reduced while the backward
            # continues (time_embed comes last: bucket 0 goes out after the backward).
            _, d_ctx = self.trainer.backward(tape, d_eps, sink=sink, need_dcontext=need_dcond,
                                             on_block_done=lambda p: self._allreduce_ready(self._final_from[p], pending))
            self._allreduce_ready(0, pending)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            _, d_ctx = self.trainer.backward(tape, d_eps, sink=sink, need_dcontext=need_dcond)
        return self._clip_and_update(d_ctx)

    def _zero_grads(self) -> None:
        # the packed slots were cleared by the previous cs_adamw_repack (and start zeroed): only the plain region needs a memset
        (self.flat_g[:self.n_plain] if self.fused else self.flat_g).zero_()

    def _step_without_objects(self, cond):
        """This rank holds no object of the global batch: zero gradients into every bucket's all-reduce, same update."""
        self.loss.zero_()
        self._zero_grads()
        if self.world > 1:
            self._allreduce_ready(0, list(range(len(self.buckets))))
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        return self._clip_and_update(torch.zeros_like(cond))

    def _clip_and_update(self, d_ctx):
        self.step_count += 1
        self.step_dev.add_(1)
        self.sumsq.zero_()
        ops_bwd.sumsq(self.flat_g, self.sumsq)
        # gradients were summed over ranks: the mean (DDP semantics) is a scale folded into the optimizer kernel
        gscale = 1.0 / self.world
        n = self.n_plain if self.fused else self.flat_p.numel()
        ops_bwd.adamw_step(self.flat_p[:n], self.flat_g[:n], self.flat_m[:n], self.flat_v[:n], lr=self.lr, betas=self.betas,
                           eps=self.eps, weight_decay=self.weight_decay, sumsq_buf=self.sumsq, max_norm=self.max_grad_norm,
                           grad_scale=gscale, step_dev=self.step_dev)
        if self.fused:
            _lib.check(_lib.load().cs_adamw_repack(
                self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                self._repack_table.data_ptr(), self._repack_n, self._repack_tiles, self._repack_taps, self._repack_tile_ci, self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, 0, self.sumsq.data_ptr(), self.max_grad_norm, gscale,
                self.step_dev.data_ptr(), torch.cuda.current_stream().cuda_stream), "cs_adamw_repack")
        self.unet._packed = None        # weights changed in place (kernel write: no autograd version bump)
        return self.loss, d_ctx

    # ------------------------------------------------------------------------------------------
    def optimizer_state_dict(self) -> dict:
        """torch.optim.AdamW-compatible state (load it with `torch.optim.AdamW(df.parameters()).load_state_dict`)."""
        hp = dict(lr=self.lr, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay)
        return flat_optimizer_state_dict(self.params, self.offsets, self.flat_m, self.flat_v, self.step_count, hp)

    def load_optimizer_state_dict(self, sd: dict) -> None:
        """Resume from a torch.optim.AdamW state dict (e.g. the 'opt' entry of a reference df_*.pth checkpoint)."""
        self.step_count = load_flat_optimizer_state(sd, self.params, self.offsets, self.flat_m, self.flat_v)
        self.step_dev.fill_(self.step_count)
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]

    # ------------------------------------------------------------------------------------------
    def capture(self, batch: int, context_dim: int, need_dcond: bool = False, warmup: int = 2) -> None:
        """Capture the whole iteration (weight re-pack, forward, backward, all-reduce, clip, AdamW) into ONE CUDA graph for
        a fixed batch size: ~1200 launches per step are then replayed without host work.  The warm-up iterations needed
        before capture run on zeros and are rolled back (parameters, moments and step counter are restored).  lr / betas /
        weight decay are baked into the graph as kernel arguments: re-capture after changing them (an LR schedule that
        changes every step should use step(), or a piecewise-constant schedule with one capture per plateau)."""
        dev = self.flat_p.device
        zs = self.model.z_shape if hasattr(self.model, "z_shape") else (3, 16, 16, 16)
        self._gz = torch.zeros((batch,) + tuple(zs), dtype=torch.float32, device=dev)
        self._gc = torch.zeros((batch, 1, context_dim), dtype=torch.float32, device=dev)
        self._gt = torch.zeros((batch,), dtype=torch.int64, device=dev)
        self._gn = torch.zeros_like(self._gz)
        saved = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.step_dev.clone(), self.step_count)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self._gz, self._gc, t=self._gt, noise=self._gn, need_dcond=need_dcond)
        torch.cuda.current_stream().wait_stream(side)

        def restore():
            self.flat_p.copy_(saved[0]); self.flat_m.copy_(saved[1]); self.flat_v.copy_(saved[2]); self.step_dev.copy_(saved[3])
            self.step_count = saved[4]
        restore()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._gout = self.step(self._gz, self._gc, t=self._gt, noise=self._gn, need_dcond=need_dcond)
        restore()      # capture does not execute, but step() bumped the host-side counter
        self.refresh_packs()        # the warm-up updates rewrote the fused-update packs: rebuild them from the restored weights
        self.graph = g

    def step_graphed(self, z: torch.Tensor, cond: torch.Tensor, t: Optional[torch.Tensor] = None,
                     noise: Optional[torch.Tensor] = None):
        """step() through the captured graph (call capture() first; same batch size).  Returns (loss, d_cond) tensors that
        are overwritten by the next replay."""
        if self.graph is None:
            raise RuntimeError("DenoiserTrainStep.step_graphed: call capture() first")
        self._check_packs_current()
        self._gz.copy_(z)
        self._gc.copy_(cond)
        if t is None:
            torch.randint(0, self.model.num_timesteps, self._gt.shape, device=self._gt.device, out=self._gt)
        else:
            self._gt.copy_(t)
        if noise is None:
            self._gn.normal_()
        else:
            self._gn.copy_(noise)
        self.graph.replay()
        self.step_count += 1
        # the replay re-packed and then updated the weights in place (raw-pointer kernels: no autograd version bump), so
        # the module's cached kernel-layout copies (emb / cross-attention matrices, DDIM graph) are stale for anyone
        # evaluating between steps: invalidate, as step() does
        self.unet._packed = None
        return self._gout


class _FlatGroup:
    """A parameter group re-homed into flat fp32 buffers (values, gradients, AdamW moments) so that the clip norm and the
    optimizer are ONE launch each and a multi-GPU step is ONE all-reduce.  The nn.Parameters become views: state_dict()
    / load_state_dict() keep working and keep the reference's keys."""

    def __init__(self, params: List[nn.Parameter]):
        self.params = params
        dev = params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]
        total = sum(sizes)
        self.flat_p, self.flat_g, self.flat_m, self.flat_v = (torch.zeros(total, dtype=torch.float32, device=dev) for _ in range(4))
        self.views: Dict[nn.Parameter, torch.Tensor] = {}
        self.offsets: Dict[nn.Parameter, int] = {}
        off = 0
        with torch.no_grad():
            for p, n in zip(params, sizes):
                v = self.flat_p[off:off + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.views[p] = self.flat_g[off:off + p.numel()].view(p.shape)
                self.offsets[p] = off
                off += n
        self.sumsq = torch.zeros((), dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros((), dtype=torch.int32, device=dev)

    def optimizer_state_dict(self, hp: dict) -> dict:
        return flat_optimizer_state_dict(self.params, self.offsets, self.flat_m, self.flat_v, int(self.step_dev.item()), hp)

    def load_optimizer_state_dict(self, sd: dict) -> None:
        self.step_dev.fill_(load_flat_optimizer_state(sd, self.params, self.offsets, self.flat_m, self.flat_v))

    def clip_and_step(self, lr, betas, eps, weight_decay, max_norm, grad_scale):
        self.step_dev.add_(1)
        self.sumsq.zero_()
        ops_bwd.sumsq(self.flat_g, self.sumsq)
        ops_bwd.adamw_step(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                           sumsq_buf=self.sumsq, max_norm=max_norm, grad_scale=grad_scale, step_dev=self.step_dev)


class ShapeBranchTrainStep:
    """One v2_full shape-branch training iteration (BASELINE cfg3 / cfg4; reference: Sg2ScVAEModel.forward :511-521 +
    train_3dfront.py:387-418): scene-graph batch -> encoder_2 (GCN-E2 + rel_mlp, BatchNorm in train mode) -> frozen VQ-VAE
    encode of the selected objects' SDFs -> denoiser forward / backward (DenoiserTrainStep) -> d_c back through rel_mlp /
    GCN-E2 / the embeddings -> clip 5.0 per group (train_3dfront.py:396-399) -> AdamW on both groups.

    Multi-GPU: every rank runs encoder_2 on the WHOLE graph batch (so BatchNorm statistics equal the
    single-process ones) and the denoiser on its own block of the selected objects; the denoiser gradients are all-reduced
    in buckets during its backward, the graph-side gradients (each rank holds the part that flows through its objects)
    in one all-reduce; both are means over ranks (DistributedDataParallel semantics)."""

    def __init__(self, model, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.01,
                 max_grad_norm: float = 5.0, loss_scale: float = 100.0, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.hp = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.max_grad_norm = max_grad_norm
        self.denoiser = DenoiserTrainStep(model.Diff, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                          max_grad_norm=max_grad_norm, loss_scale=loss_scale, group=group)
        self.graph_params = _FlatGroup([p for p in model._enc2_params() if p.requires_grad])

    def shard(self, n_selected: int):
        """This rank's contiguous block of the selected objects (parallel.py's block partition)."""
        from .parallel import partition
        return partition(n_selected, self.world)[self.rank]

    def step(self, z, objs, triples, text_feat, rel_feat, sdfs, rows: Optional[torch.Tensor] = None,
             t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None, n_total: Optional[int] = None):
        """z (O, 64) layout latents, objs (O,), triples (T, 3), CLIP features (O, 512) / (T, 512), sdfs (O, 1, 64, 64, 64).
        rows: indices of the objects that enter the denoiser on THIS rank (default: this rank's block of all O objects;
        may be empty); n_total: objects entering the denoiser over all ranks (default: O, or len(rows) * world for given rows);
        t / noise: optional fixed timesteps and noise for those rows.  Returns (loss [mean MSE on this rank], d_z (O, 64))."""
        m = self.model
        O = objs.shape[0]
        if rows is None:
            lo, hi = self.shard(O)
            rows = torch.arange(lo, hi, device=objs.device)
            n_total = O
        elif n_total is None:       # caller-chosen rows: the global count is the sum over ranks
            n_total = int(rows.shape[0]) * self.world
        n_local = int(rows.shape[0])
        # global-batch mean (not a mean of per-rank means): weight this rank's mean gradient by n_local * world / n_total
        weight = n_local * self.world / max(n_total, 1)
        uc, c, tape = m.encoder_2_train(z, objs, triples, text_feat, rel_feat)
        c = uc if c is None else c
        if n_local:
            with torch.no_grad():
                lat = m.Diff.vqvae(sdfs[rows].to(c.device), forward_no_quant=True, encode_only=True)
        else:
            lat = torch.zeros((0,) + tuple(m.Diff.z_shape), dtype=torch.float32, device=c.device)
        loss, d_c_rows = self.denoiser.step(lat, c[rows].contiguous(), t=t, noise=noise, need_dcond=True, grad_weight=weight)
        d_c = torch.zeros((O, c.shape[-1]), dtype=torch.float32, device=c.device)
        if n_local:
            d_c.index_copy_(0, rows, d_c_rows.reshape(n_local, -1).float())
        g = self.graph_params
        g.flat_g.zero_()
        _, d_z = m.encoder_2_backward(tape, d_c=d_c if m.use_E2 else None, d_uc=None if m.use_E2 else d_c, sink=GradSink(g.views))
        if self.world > 1:
            dist.all_reduce(g.flat_g, op=dist.ReduceOp.SUM, group=self.group)
        g.clip_and_step(max_norm=self.max_grad_norm, grad_scale=1.0 / self.world, **self.hp)
        return loss, d_z

    # ------------------------------------------------------------------------------------------
    def capture(self, n_objs: int, n_triples: int, n_rows: Optional[int] = None, z_dim: int = 64, clip_dim: int = 512,
                sdf_res: int = 64, warmup: int = 2) -> None:
        """Capture the WHOLE iteration -- encoder_2 forward, frozen VQ-VAE encode, denoiser re-pack / forward / backward with
        its bucketed all-reduce, the graph-side backward and its all-reduce, clip + AdamW of both groups -- into ONE CUDA
        graph for a fixed batch geometry (n_objs objects, n_triples triples, n_rows of them entering the denoiser on this
        rank; default: this rank's block).  Eager, the iteration is host-launch bound (~2300 launches; 107 ms against 82 ms
        for the graphed denoiser part alone on 4 GPUs).  Inputs are copied into static buffers by step_graphed().  The
        warm-up iterations run on zeros and are rolled back: both flat parameter groups, their moments and step counters,
        and every module buffer (BatchNorm running statistics)."""
        m = self.model
        dev = self.graph_params.flat_p.device
        if n_rows is None:
            lo, hi = self.shard(n_objs)
            rows = torch.arange(lo, hi, device=dev)
            n_total = n_objs
        else:
            rows = torch.arange(n_rows, device=dev)
            n_total = n_rows * self.world
        zs = tuple(m.Diff.z_shape)
        self._s = dict(z=torch.zeros(n_objs, z_dim, device=dev), objs=torch.zeros(n_objs, dtype=torch.int64, device=dev),
                       triples=torch.zeros(n_triples, 3, dtype=torch.int64, device=dev),
                       text=torch.zeros(n_objs, clip_dim, device=dev), rel=torch.zeros(n_triples, clip_dim, device=dev),
                       sdfs=torch.zeros(n_objs, 1, sdf_res, sdf_res, sdf_res, device=dev), rows=rows,
                       t=torch.zeros(rows.shape[0], dtype=torch.int64, device=dev),
                       noise=torch.zeros((rows.shape[0],) + zs, device=dev))
        self._s_total = n_total
        d, gp = self.denoiser, self.graph_params
        saved = [t.clone() for t in (d.flat_p, d.flat_m, d.flat_v, d.step_dev, gp.flat_p, gp.flat_m, gp.flat_v, gp.step_dev)]
        saved_count = d.step_count
        bufs = {n: b.clone() for n, b in m.named_buffers()}

        def restore():
            for dst, src in zip((d.flat_p, d.flat_m, d.flat_v, d.step_dev, gp.flat_p, gp.flat_m, gp.flat_v, gp.step_dev), saved):
                dst.copy_(src)
            d.step_count = saved_count
            with torch.no_grad():
                for n, b in m.named_buffers():
                    b.copy_(bufs[n])

        def run():
            st = self._s
            return self.step(st["z"], st["objs"], st["triples"], st["text"], st["rel"], st["sdfs"], rows=st["rows"], t=st["t"],
                             noise=st["noise"], n_total=self._s_total)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                run()
        torch.cuda.current_stream().wait_stream(side)
        restore()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._gout = run()
        restore()                       # capture does not execute, but step() bumped the host-side counters
        d.refresh_packs()               # (the warm-up updates rewrote the fused-update packs)
        self.graph = g

    def step_graphed(self, z, objs, triples, text_feat, rel_feat, sdfs, rows: Optional[torch.Tensor] = None,
                     t: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None):
        """step() through the captured graph (capture() first, same geometry).  rows: which objects enter the denoiser on this
        rank (same count as captured; default: the captured rows).  Returns (loss, d_z): overwritten by the next replay."""
        if getattr(self, "graph", None) is None:
            raise RuntimeError("ShapeBranchTrainStep.step_graphed: call capture() first")
        self.denoiser._check_packs_current()
        st = self._s
        for key, src in (("z", z), ("objs", objs), ("triples", triples), ("text", text_feat), ("rel", rel_feat), ("sdfs", sdfs)):
            if tuple(src.shape) != tuple(st[key].shape):
                raise ValueError(f"step_graphed: {key} has shape {tuple(src.shape)}, captured {tuple(st[key].shape)}")
            st[key].copy_(src)
        if rows is not None:
            if tuple(rows.shape) != tuple(st["rows"].shape):
                raise ValueError("step_graphed: the number of denoiser rows is part of the captured geometry")
            st["rows"].copy_(rows)
        if t is None:
            torch.randint(0, self.model.Diff.num_timesteps, st["t"].shape, device=st["t"].device, out=st["t"])
        else:
            st["t"].copy_(t)
        if noise is None:
            st["noise"].normal_()
        else:
            st["noise"].copy_(noise)
        self.graph.replay()
        self.denoiser.step_count += 1
        self.denoiser.unet._packed = None       # weights were updated in place by raw-pointer kernels (see DenoiserTrainStep)
        return self._gout
