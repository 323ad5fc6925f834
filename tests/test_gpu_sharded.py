"""GPU parity of the multi-GPU row-stripe path (imagepipe_b200/sharded.py) and of the banded host<->device
path.  The stripe driver is exercised on ONE GPU by walking the per-rank layouts in turn (the kernels and the
layouts are exactly those of an N-rank run; the exchange itself is covered by the gloo tests on CPU and by
test_sharded_nccl below when the box has 2+ GPUs)."""
import os
import socket

import numpy as np
import pytest

import common
from common import assert_bit_exact

pytestmark = pytest.mark.gpu


def _pipeline_for(ip, ctx, width, height, params, settings_kw):
    from imagepipe_b200 import _capi
    dummy = ip.DeviceArray(64, ctx)  # never read: every launch gets a stripe source
    src = ip.ImageSource(_capi.SRC_RAW_U16, width, height, 1, dummy)
    p = ip.Pipeline.new_from_source(src, ctx=ctx)
    common.fill_ipb_ops(p.ops, params)
    for k, v in settings_kw.items():
        setattr(p.globals.settings, k, v)
    return p


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("case", ["full", "scaled", "xtrans"])
def test_stripe_layouts_on_one_gpu(ip, orc, ctx, world, case):
    from imagepipe_b200.sharded import plan_stripes, run_stripe_8bit
    w, h = 704, 520
    cfa = common.XTRANS if case == "xtrans" else "RGGB"
    st = {"maxwidth": 176} if case == "scaled" else {}
    params = common.raw_params(cfa=cfa)
    data = common.synth_cfa(w, h, seed=77)
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params, st))
    p = _pipeline_for(ip, ctx, w, h, params, st)
    lays = plan_stripes(p.ops, p.globals.settings, w, h, world)
    parts = []
    for lay in lays:
        rows = ip.DeviceArray.from_numpy(data[lay.src_row0:lay.src_row1], ctx)
        dst = ip.DeviceArray((lay.out_row1 - lay.out_row0) * lay.out_width * 3, ctx)
        run_stripe_8bit(p, rows.ptr, lay, dst)
        parts.append(dst.to_numpy(np.uint8, (lay.out_row1 - lay.out_row0, lay.out_width, 3)))
    assert_bit_exact(np.concatenate(parts, 0), want, f"{case} x{world} stripes vs oracle")


@pytest.mark.parametrize("case", ["full", "scaled", "crop"])
def test_banded_host_path_equals_oracle(ip, orc, ctx, case):
    """Host source + host destination large enough for several bands (H2D / kernel / D2H overlapped)."""
    w, h = 4096, 1500
    data = common.synth_cfa(w, h, seed=91)
    st = {"maxwidth": 1024} if case == "scaled" else {}
    params = common.raw_params(crops=(3, 8, 5, 16) if case == "crop" else (0, 0, 0, 0))
    want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", params, st))
    n0 = ctx.launch_count
    pg = common.make_ipb_pipeline(ip, data, "raw", params, st, ctx=ctx)
    pg.set_band_mb(4)
    got = pg.output_8bit()
    assert ctx.launch_count - n0 > 1  # really ran in bands
    assert_bit_exact(got.to_numpy(), want, f"banded {case}")
    # device-resident source, host destination: kernel / D2H overlap only
    pg = common.make_ipb_pipeline(ip, data, "raw", params, st, ctx=ctx, on_device=True)
    pg.set_band_mb(2)
    assert_bit_exact(pg.output_8bit().to_numpy(), want, f"banded {case}, device source")
    # one band (whole-frame copies)
    pg = common.make_ipb_pipeline(ip, data, "raw", params, st, ctx=ctx)
    pg.set_band_mb(0)
    n0 = ctx.launch_count
    assert_bit_exact(pg.output_8bit().to_numpy(), want, f"unbanded {case}")
    assert ctx.launch_count - n0 == 1


def test_banded_16bit(ip, orc, ctx):
    w, h = 3008, 1400
    data = common.smooth_cfa(w, h)
    params = common.raw_params(cfa="GBRG")
    want = orc.pipeline_output_16bit(orc.make_pipeline(data, "raw", params))
    pg = common.make_ipb_pipeline(ip, data, "raw", params, ctx=ctx)
    pg.set_band_mb(4)
    assert_bit_exact(pg.output_16bit().to_numpy(), want, "banded output_16bit")


# ---------------------------------------------------------------------------------------------- 2+ GPUs

def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _nccl_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import imagepipe_b200 as ip
    from imagepipe_b200.sharded import DevicePtr, exchange_halos, plan_stripes, run_stripe_8bit
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        w, h = 1920, 1088
        params = common.raw_params()
        stream = torch.cuda.Stream()
        ctx = ip.Context(rank, stream.cuda_stream)
        p = _pipeline_for(ip, ctx, w, h, params, {})
        lays = plan_stripes(p.ops, p.globals.settings, w, h, world)
        me = lays[rank]
        nframes = 3   # several frames in flight: their halo rows travel packed, one message per neighbour
        with torch.cuda.stream(stream):
            bufs = [torch.full((me.src_row1 - me.src_row0, w), 0x7FFF, dtype=torch.int16, device="cuda") for _ in range(nframes)]
            for k, buf in enumerate(bufs):
                own = buf[me.own_row0 - me.src_row0: me.own_row1 - me.src_row0]
                # every rank generates only its own block of each frame, on its own GPU
                ip.lib().ipb_synth_cfa_u16(ctx.handle, common.SEED + k, w, me.own_row0, me.own_row1 - me.own_row0, own.data_ptr())
            # the exchange goes through the C ABI (ipb_halo_exchange: NCCL send/recv between stripe neighbours);
            # torch.distributed only hands the 128-byte communicator id to the ranks
            from imagepipe_b200.sharded import Comm, exchange_halos_nccl
            box = [Comm.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            comm = Comm(box[0], rank, world, rank, stream.cuda_stream)
            exchange_halos_nccl(comm, [bufs[0].data_ptr()], lays, w * 2)            # one buffer: a message per row block
            exchange_halos_nccl(comm, [b.data_ptr() for b in bufs], lays, w * 2)    # all of them: packed
            outs = []
            for buf in bufs:
                out = torch.empty(((me.out_row1 - me.out_row0), me.out_width, 3), dtype=torch.uint8, device="cuda")
                run_stripe_8bit(p, buf.data_ptr(), me, DevicePtr(out.data_ptr(), out.numel(), out))
                outs.append(out)
        stream.synchronize()
        comm.close()
        images = []
        for out in outs:
            gathered = [torch.empty((l.out_row1 - l.out_row0, l.out_width, 3), dtype=torch.uint8, device="cuda") for l in lays]
            if rank == 0:  # stripes may differ in height: gather through rank 0 with send/recv
                gathered[0] = out
                for r in range(1, world):
                    dist.recv(gathered[r], r)
                images.append(torch.cat(gathered, 0).cpu().numpy())
            else:
                dist.send(out, 0)
        if rank == 0:
            ret["images"] = images
    finally:
        dist.destroy_process_group()


def test_sharded_nccl(orc):
    import torch
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs 2+ GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for k, image in enumerate(ret["images"]):
        data = common.synth_cfa(1920, 1088, seed=common.SEED + k)
        want = orc.pipeline_output_8bit(orc.make_pipeline(data, "raw", common.raw_params()))
        assert_bit_exact(image, want, f"{world}-GPU NCCL halo exchange + stripes vs oracle, frame {k}")
