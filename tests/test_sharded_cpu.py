"""Host logic of the multi-GPU row-stripe path (imagepipe_b200/sharded.py) on CPU: partitioning, the source rows
each stripe needs (ipb_stripe_plan: pure host arithmetic), and the halo exchange over torch.distributed with the
gloo backend at world sizes 2 and 3.  After the exchange every rank must hold exactly the frame rows its stripe's
kernel reads; the kernel itself is covered by the GPU suite (tests/test_gpu_fused.py, tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest

import common


def make_ops(ip, cfa="RGGB", crops=(0, 0, 0, 0)):
    src = ip.ImageSource.Raw(np.zeros((16, 16), np.uint16))
    ops = ip.PipelineOps.new(src)
    common.fill_ipb_ops(ops, common.raw_params(cfa=cfa, crops=crops))
    return ops


def settings(ip, **kw):
    s = ip.PipelineSettings.default()
    for k, v in kw.items():
        setattr(s, k, v)
    return s


def test_partition_rows():
    from imagepipe_b200.sharded import partition_rows
    for n, w in [(8736, 8), (4000, 3), (100, 8), (31, 2), (1, 4)]:
        parts = partition_rows(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n and len(parts) == w
        for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
            assert a1 == b0 and a0 <= a1
        if n // w >= 2:
            assert all(a % 2 == 0 for a, _ in parts[1:])       # boundaries on the Bayer period
        if n // w >= 32:
            assert all(a % 32 == 0 for a, _ in partition_rows(n, w, 32)[1:])
        if n >= w:
            assert all(b > a for a, b in parts)
    assert partition_rows(8736, 8)[0] == (0, 1092)             # C5: eight equal stripes
    assert partition_rows(8736, 8, 32)[0] == (0, 1088)         # ... or rounded to the tile height


def test_stripe_plan_full_resolution(ip):
    from imagepipe_b200.sharded import stripe_plan
    ops = make_ops(ip)
    assert stripe_plan(ops, None, 11648, 8736) [2:] == (11648, 8736)
    assert stripe_plan(ops, None, 11648, 8736, 0, 1088)[:2] == (0, 1089)          # top stripe: one row below
    assert stripe_plan(ops, None, 11648, 8736, 1088, 2176)[:2] == (1087, 2177)     # interior: one above, one below
    assert stripe_plan(ops, None, 11648, 8736, 8000, 8736)[:2] == (7999, 8736)     # bottom stripe
    cropped = make_ops(ip, crops=(4, 0, 2, 0))  # top 4, bottom 2
    assert stripe_plan(cropped, None, 600, 400)[2:] == (600, 394)
    assert stripe_plan(cropped, None, 600, 400, 32, 64)[:2] == (35, 69)            # shifted by the top crop


def test_stripe_plan_scaled(ip):
    from imagepipe_b200.sharded import stripe_plan
    ops = make_ops(ip)
    st = settings(ip, maxwidth=1500, maxheight=1000)
    assert stripe_plan(ops, st, 6000, 4000)[2:] == (1500, 1000)
    s0, s1, _, _ = stripe_plan(ops, st, 6000, 4000, 512, 768)
    # scaling.rs:72,86-87: skip = 3999/999; rows floor(skip*512) .. floor(skip*768)
    skip = np.float32(3999.0) / np.float32(999.0)
    assert s0 == int(np.floor(skip * np.float32(512))) and s1 == int(np.floor(skip * np.float32(768))) + 1


def test_stripe_plan_rejects_unshardable_chains(ip):
    from imagepipe_b200.sharded import stripe_plan
    ops = make_ops(ip)
    ops.transform.rotation = 1  # Rotate90: row stripes would become column stripes
    with pytest.raises(ip.IpbError):
        stripe_plan(ops, None, 600, 400)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("case", ["full", "scaled", "xtrans"])
def test_plan_stripes_covers_the_frame(ip, world, case):
    from imagepipe_b200.sharded import plan_stripes
    w, h = (1200, 1000)
    ops = make_ops(ip, cfa=common.XTRANS if case == "xtrans" else "RGGB")
    st = settings(ip, maxwidth=300) if case == "scaled" else None
    lays = plan_stripes(ops, st, w, h, world)
    assert [l.rank for l in lays] == list(range(world))
    assert lays[0].out_row0 == 0 and lays[-1].out_row1 == lays[0].out_height
    assert lays[0].own_row0 == 0 and lays[-1].own_row1 == h
    for a, b in zip(lays, lays[1:]):
        assert a.out_row1 == b.out_row0 and a.own_row1 == b.own_row0
    for l in lays:
        assert l.src_row0 <= l.own_row0 <= l.own_row1 <= l.src_row1
        if l.out_row1 > l.out_row0 and case == "full":
            assert l.src_row0 == max(l.out_row0 - 1, 0) and l.src_row1 == min(l.out_row1 + 1, h)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _halo_worker(rank, world, port, case, ret):
    import torch
    import torch.distributed as dist
    import imagepipe_b200 as ip
    from imagepipe_b200.sharded import exchange_halos, plan_stripes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w, h = 640, 480
        ops = make_ops(ip)
        st = settings(ip, maxwidth=160) if case == "scaled" else None
        lays = plan_stripes(ops, st, w, h, world)
        me = lays[rank]
        frame = common.synth_cfa(w, h)
        # each rank starts with its own rows only (as if decoded / generated there); halos are poisoned
        buf = torch.full((me.src_row1 - me.src_row0, w), 0x7FFF, dtype=torch.int16)
        own = torch.from_numpy(frame[me.own_row0:me.own_row1].view(np.int16).copy())
        buf[me.own_row0 - me.src_row0: me.own_row1 - me.src_row0] = own
        # a second frame with the same layout travels in the same batched group
        frame2 = common.synth_cfa(w, h, seed=common.SEED + 9)
        buf2 = torch.full((me.src_row1 - me.src_row0, w), 0x7FFF, dtype=torch.int16)
        buf2[me.own_row0 - me.src_row0: me.own_row1 - me.src_row0] = torch.from_numpy(
            frame2[me.own_row0:me.own_row1].view(np.int16).copy())
        exchange_halos([buf, buf2], lays, rank)
        got = buf.numpy().view(np.uint16)
        ok = bool(np.array_equal(got, frame[me.src_row0:me.src_row1])) and \
            bool(np.array_equal(buf2.numpy().view(np.uint16), frame2[me.src_row0:me.src_row1]))
        halo_rows = (me.own_row0 - me.src_row0) + (me.src_row1 - me.own_row1)
        flags = torch.tensor([int(ok), halo_rows])
        gathered = [torch.zeros_like(flags) for _ in range(world)]
        dist.all_gather(gathered, flags)
        if rank == 0:
            ret["ok"] = [int(g[0]) for g in gathered]
            ret["halo"] = [int(g[1]) for g in gathered]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["full", "scaled"])
def test_halo_exchange_gloo(world, case):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_halo_worker, args=(world, _free_port(), case, ret), nprocs=world, join=True)
    assert ret["ok"] == [1] * world
    if case == "full":  # one raw row per neighbour (3x3 stencil): edge ranks 1, interior ranks 2
        assert ret["halo"] == [1] + [2] * (world - 2) + [1]
    else:  # adjacent windows share exactly one source row (to_y of row r-1 == from_y of row r, scaling.rs:86-87)
        assert sum(ret["halo"]) == world - 1
