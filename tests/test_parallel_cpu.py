"""CPU suite, part 3: the multi-GPU host logic (object sharding, ragged all_gather) under a world_size-2 gloo group."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from commonscenes_b200 import parallel


def test_partition_covers_everything_in_order():
    for n in (0, 1, 7, 10, 32, 33):
        for w in (1, 2, 4, 8):
            b = parallel.partition(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert parallel.partition(10, 8) == [(0, 2), (2, 4), (4, 5), (5, 6), (6, 7), (7, 8), (8, 9), (9, 10)]   # cfg5: 2,2,1,1,1,1,1,1


def test_cfg_pair_units_balance_cfg5():
    """cfg5: 10 objects x CFG = 20 forwards over 8 ranks -> 3,3,3,3,2,2,2,2 instead of the 2:1 object split."""
    sizes = [hi - lo for lo, hi in parallel.pair_units(10, 8)]
    assert sizes == [3, 3, 3, 3, 2, 2, 2, 2] and parallel.pair_units(10, 8)[-1][1] == 20
    assert [hi - lo for lo, hi in parallel.pair_units(1, 2)] == [1, 1]      # one object: uncond on rank 0, cond on rank 1


class _StubDiff:
    """Stands in for SDFusionText2ShapeModel on CPU: 'decodes' each object to a volume filled with a value that
    depends only on that object's conditioning, so the gathered result reveals any mis-ordering."""
    z_shape = (3, 2, 2, 2)

    def rel2shape(self, data, ddim_steps=100, ddim_eta=0.0, uc_scale=3.0, seed=None):
        v = data["rel"].reshape(data["rel"].shape[0], -1).sum(dim=1) + 0.5 * data["uc"].reshape(data["uc"].shape[0], -1).sum(dim=1)
        return v.view(-1, 1, 1, 1, 1).expand(-1, 1, 8, 8, 8).contiguous()


def _worker(rank, world, port, n_obj, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        data = {"sdf": torch.zeros(n_obj, 1), "rel": torch.randn(n_obj, 1, 4, generator=g), "uc": torch.randn(n_obj, 1, 4, generator=g)}
        out = parallel.rel2shape_sharded(_StubDiff(), data, seed=1)
        ref = _StubDiff().rel2shape(data)
        ok = out.shape == ref.shape and torch.equal(out, ref)
        loc = parallel.shard(data["rel"], rank, world)
        back = parallel.gather_objects(loc, n_obj)
        ok = ok and torch.equal(back, data["rel"])
        # CFG-pair split: each rank evaluates its block of the 2 * O [uncond; cond] forwards, exchange_eps rebuilds the batch
        full = torch.randn(2 * n_obj, 3, 2, 2, 2, generator=torch.Generator().manual_seed(7))
        ulo, uhi = parallel.pair_units(n_obj, world)[rank]
        ok = ok and torch.equal(parallel.exchange_eps(full[ulo:uhi].clone(), n_obj), full)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_sampling_gathers_objects_in_order_world2():
    for n_obj in (5, 1):          # ragged split (3 + 2) and more ranks than objects (1 + 0)
        ctx = mp.get_context("spawn")
        ret = ctx.Manager().dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_obj, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}


# ---- training partition: bucketed gradient all-reduce scheduling of DenoiserTrainStep ----
def test_gradient_buckets_cover_the_flat_buffer():
    from commonscenes_b200.train import make_buckets
    sizes = [4, 100, 28, 64, 8, 300, 12]
    b = make_buckets(sizes, 128)
    assert b[0][0] == 0 and b[-1][1] == sum(sizes) and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    assert b == [(0, 132), (132, 504), (504, 516)]
    assert make_buckets([8, 8], 1 << 30) == [(0, 16)] and make_buckets([], 4) == []


def _train_worker(rank, world, port, ret):
    from commonscenes_b200.train import DenoiserTrainStep, make_buckets
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [16, 48, 32, 8, 64, 24]
        offs = [sum(sizes[:i]) for i in range(len(sizes))]
        st = object.__new__(DenoiserTrainStep)          # only the scheduling state: no CUDA in this container
        st.buckets, st.comm_stream, st.group = make_buckets(sizes, 60), None, None
        g = torch.Generator().manual_seed(rank)
        st.flat_g = torch.randn(sum(sizes), generator=g)
        mine = st.flat_g.clone()
        other = torch.randn(sum(sizes), generator=torch.Generator().manual_seed(1 - rank))
        pending = list(range(len(st.buckets)))
        reduced_after = []
        for off in reversed(offs):                      # the backward finishes parameters from the last to the first
            st._allreduce_ready(off, pending)
            reduced_after.append(len(st.buckets) - len(pending))
        st._allreduce_ready(0, pending)
        ok = not pending and torch.allclose(st.flat_g, mine + other)           # every bucket reduced exactly once
        ok = ok and reduced_after == sorted(reduced_after) and reduced_after[0] <= 1   # buckets go out as soon as they are final
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_schedule_world2():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert dict(ret) == {0: True, 1: True}


def test_fused_gradient_layout_keeps_the_bucket_schedule_valid():
    """Fused-update mode (packed-gradient slots behind the plain parameters): walking the parameters from the last to the
    first, `final_from` never decreases below data that is still being written, and every bucket goes out exactly once."""
    from commonscenes_b200.train import fused_gradient_layout, make_buckets
    sizes = [8, 400, 16, 16, 1200, 8, 640, 32, 24]
    slots = [None, 448, None, None, 1280, None, 704, None, None]      # conv weights own (larger, padded) packed slots
    p_off, g_off, final_from, n_plain, order = fused_gradient_layout(sizes, slots)
    assert n_plain == sum(s for s, k in zip(sizes, slots) if k is None)
    assert sorted(order) == list(range(len(sizes))) and [i for i in order if slots[i] is None] == [0, 2, 3, 5, 7, 8]
    total_g = n_plain + sum(k for k in slots if k is not None)
    # plain parameters keep identical offsets in flat_p and flat_g; packed slots follow in module order
    assert all(p_off[i] == g_off[i] for i in range(len(sizes)) if slots[i] is None)
    assert [g_off[i] for i in (1, 4, 6)] == [n_plain, n_plain + 448, n_plain + 448 + 1280]
    buckets = make_buckets([sizes[i] if slots[i] is None else slots[i] for i in order], 600)
    assert buckets[0][0] == 0 and buckets[-1][1] == total_g
    written_from = total_g                     # lowest flat_g offset a not-yet-visited parameter may still write to
    sent = set()
    for i in reversed(range(len(sizes))):      # the backward visits parameters from the last to the first
        lo = final_from[i]
        # everything at or after `lo` must belong to parameters already visited (index >= i)
        for j in range(i):
            if slots[j] is not None:
                assert g_off[j] + slots[j] <= lo
            else:
                assert g_off[j] + sizes[j] <= n_plain <= lo
        for b, (blo, bhi) in enumerate(buckets):
            if blo >= lo:
                sent.add(b)
        assert lo <= written_from
        written_from = lo
    assert final_from[0] == n_plain            # the plain region is only final after the whole backward
    sent |= set(range(len(buckets)))           # the final _allreduce_ready(0, ...) sends the rest
    assert sent == set(range(len(buckets)))
