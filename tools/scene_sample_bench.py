"""BASELINE cfg5: full ancestral sampling of ONE scene (10 objects, livingroom-sized graph) with classifier-free guidance,
1000 DDPM steps, VQ-VAE decode to 64^3 SDFs; objects sharded over the ranks (2,2,1,1,1,1,1,1 at 8 GPUs), one all_gather
of the SDFs at the end.  Also times the reference's own sampling mode (DDIM, S = 100) on the same scene.

    python tools/scene_sample_bench.py                       # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/scene_sample_bench.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from commonscenes_b200 import parallel
from commonscenes_b200.model.sdfusion_txt2shape_model import SDFusionText2ShapeModel, default_opt

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(111)
m = SDFusionText2ShapeModel(default_opt(device=f"cuda:{local}"))
with torch.no_grad():
    for p in m.df.parameters():
        if p.dim() > 1 and float(p.abs().max()) == 0:
            torch.nn.init.normal_(p, std=0.02)
n = 10
g = torch.Generator(device="cuda").manual_seed(5)
data = {"sdf": torch.zeros(n, 1, 64, 64, 64, device="cuda"), "rel": torch.randn(n, 1, 1280, device="cuda", generator=g),
        "uc": torch.randn(n, 1, 1280, device="cuda", generator=g)}


def run(**kw):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = parallel.rel2shape_sharded(m, data, uc_scale=3.0, seed=7, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    assert out.shape == (n, 1, 64, 64, 64) and torch.isfinite(out).all()
    return float(dt)


run(ddim_steps=5)                                   # warm-up: packing, graph capture
run(ddpm_timesteps=20, sampler="ddpm")
t_ddim = run(ddim_steps=100)
t_ddpm = run(sampler="ddpm")
if rank == 0:
    per_rank = [hi - lo for lo, hi in parallel.partition(n, world)]
    print(f"cfg5 scene sampling, {n} objects on {world} GPU(s) (objects per rank {per_rank}): "
          f"DDIM S=100 + decode {t_ddim:.2f} s | ancestral DDPM 1000 steps + decode {t_ddpm:.2f} s "
          f"({1000 * n / t_ddpm:.0f} object-steps/s)", flush=True)
if world > 1:
    dist.destroy_process_group()
