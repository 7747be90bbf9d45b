"""Data-parallel training check + timing (BASELINE cfg4 at N ranks; run under torchrun, one rank per GPU):

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/ddp_check.py [--objs-per-rank 4] [--time]

1. correctness: the N-rank ShapeBranchTrainStep (objects block-partitioned over ranks, bucketed NCCL all-reduce of the
   denoiser gradients overlapped with the backward, one all-reduce of the graph-side gradients) must produce the gradients
   of ONE process running the whole batch: compares flat gradient buffers / world against a single-process step on the same
   (t, noise).
2. `--time`: train-steps/s at 32 objects per rank (weak scaling, cfg4), max over ranks.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_shape_branch_train_gpu import _model, _scene_batch      # synthetic cfg3 batch + seeded model (test infrastructure)
from commonscenes_b200.train import ShapeBranchTrainStep


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = int(sys.argv[sys.argv.index("--objs-per-rank") + 1]) if "--objs-per-rank" in sys.argv else 4
    O = per * world
    batch = {k: v.cuda() for k, v in _scene_batch(world, per, 6, 36, 16, seed=3).items()}
    g = torch.Generator().manual_seed(4)
    t = torch.randint(0, 1000, (O,), generator=g).cuda()
    noise = torch.randn(O, 3, 16, 16, 16, generator=g).cuda()
    args = (batch["z"], batch["objs"], batch["triples"], batch["text"], batch["rel"], batch["sdfs"])

    multi = ShapeBranchTrainStep(_model(77))
    lo, hi = multi.shard(O)
    rows = torch.arange(lo, hi, device="cuda")
    loss_m, _ = multi.step(*args, rows=rows, t=t[lo:hi], noise=noise[lo:hi])
    solo_group = [dist.new_group([r]) for r in range(world)][rank]       # a world-size-1 group: the single-process step
    solo = ShapeBranchTrainStep(_model(77), group=solo_group)
    assert solo.world == 1 and solo.denoiser.world == 1
    loss_s, _ = solo.step(*args, rows=torch.arange(O, device="cuda"), t=t, noise=noise)
    torch.cuda.synchronize()
    res = {}
    for name, a, b in (("denoiser", multi.denoiser.flat_g / world, solo.denoiser.flat_g),
                       ("graph", multi.graph_params.flat_g / world, solo.graph_params.flat_g)):
        res[name] = (float((a - b).norm() / b.norm()), float((a * b).sum() / (a.norm() * b.norm())))
    lm = loss_m.clone()
    dist.all_reduce(lm)
    ok = all(r < 3e-2 and c > 0.999 for r, c in res.values()) and abs(float(lm) / world - float(loss_s)) / float(loss_s) < 1e-2
    # all ranks must hold identical parameters after the step
    chk = torch.stack([multi.denoiser.flat_p.double().sum(), multi.graph_params.flat_p.double().sum()])
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo_, hi_))
    if rank == 0:
        print(f"ddp_check world={world}: gradient (rel-L2, cosine) vs single process: {res}; mean loss {float(lm) / world:.5f} vs {float(loss_s):.5f}; "
              f"replicas identical after the step: {same} -> {'OK' if ok and same else 'FAIL'}", flush=True)
    del solo
    if "--time" in sys.argv:
        per = 32
        O = per * world
        batch = {k: v.cuda() for k, v in _scene_batch(4 * world, 8, 12, 36, 16, seed=5).items()}
        args = (batch["z"], batch["objs"], batch["triples"], batch["text"], batch["rel"], batch["sdfs"])
        lo, hi = multi.shard(O)
        rows = torch.arange(lo, hi, device="cuda")
        for _ in range(3):
            multi.step(*args, rows=rows)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        e0.record()
        for _ in range(K):
            multi.step(*args, rows=rows)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"cfg4-style eager train step, {per} objects/rank x {world} ranks: {float(ms):.1f} ms -> {1000 / float(ms):.2f} steps/s, "
                  f"{O * 1000 / float(ms):.1f} objects/s", flush=True)
    if "--graph" in sys.argv:      # denoiser-only data-parallel step captured in ONE CUDA graph (NCCL all-reduces inside)
        den = multi.denoiser
        z = torch.randn(32, 3, 16, 16, 16, device="cuda")
        ctx = torch.randn(32, 1, 1280, device="cuda")
        den.capture(32, 1280)
        for _ in range(3):
            den.step_graphed(z, ctx)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for _ in range(K):
            den.step_graphed(z, ctx)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        chk = den.flat_p.double().sum().reshape(1)
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"graphed data-parallel denoiser step (forward, backward, bucketed NCCL all-reduce, clip, AdamW), 32 objects/rank x "
                  f"{world} ranks: {float(ms):.1f} ms -> {1000 / float(ms):.2f} steps/s, {32 * world * 1000 / float(ms):.1f} objects/s; "
                  f"replicas identical: {bool(torch.equal(lo_, hi_))}", flush=True)
    dist.destroy_process_group()
    if not (ok and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
