"""Row-stripe sharding of one large raw frame over the GPUs of one box (BASELINE config 5).

The reference has no equivalent (it is a single-process Rayon program); this is the multi-GPU form of its
`Pipeline::output_8bit` for frames whose raw samples are already spread over the devices in disjoint row blocks.
Only two steps of the path are not pointwise: demosaic::full needs one raw row above and below every stripe
(3x3 stencil, demosaic.rs:70-74) and scaled_demosaic needs the rows of its windows (scaling.rs:84-87).  Those
halo rows are the only data exchanged: one P2P send/recv pair per stripe boundary through torch.distributed
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  Everything else — which rows a stripe needs, how the
rows are partitioned — is host arithmetic done identically on every rank (ipb_stripe_plan, no GPU needed).

    layouts = plan_stripes(ops, settings, width, height, world)
    buf     = alloc rows [lay.src_row0, lay.src_row1) ; fill the owned block ; exchange_halos(buf, layouts, rank)
    out     = run_stripe_8bit(pipeline, buf_ptr, lay, dst)        # fused kernel in full-frame coordinates
"""
import ctypes as C
from dataclasses import dataclass

from . import _capi
from ._capi import lib


@dataclass(frozen=True)
class StripeLayout:
    rank: int
    out_row0: int      # output rows [out_row0, out_row1) of the full-frame result
    out_row1: int
    src_row0: int      # source rows [src_row0, src_row1) the stripe's kernel reads (own block + halos)
    src_row1: int
    own_row0: int      # source rows [own_row0, own_row1) this rank holds before the exchange (disjoint over ranks)
    own_row1: int
    out_width: int
    out_height: int

    @property
    def halo_up(self):    # rows received from rank - 1
        return self.src_row0, self.own_row0

    @property
    def halo_down(self):  # rows received from rank + 1
        return self.own_row1, self.src_row1


def partition_rows(nrows, world, multiple=2):
    """Balanced split of output rows; interior boundaries on multiples of `multiple` (2: the Bayer period — the kernels
    work in full-frame coordinates, so any boundary is correct; even ones keep a stripe's first row on the pattern's
    first row), relaxed for short frames so that no stripe is empty while nrows >= world."""
    while multiple > 1 and nrows // world < multiple:
        multiple //= 2
    bounds = [0]
    for k in range(1, world):
        b = (nrows * k // world + multiple // 2) // multiple * multiple
        bounds.append(min(max(b, bounds[-1]), nrows))
    bounds.append(nrows)
    return [(bounds[k], bounds[k + 1]) for k in range(world)]


def stripe_plan(ops, settings, width, height, out_row0=0, out_row1=0):
    """ipb_stripe_plan: (src_row0, src_row1, out_width, out_height) for output rows [out_row0, out_row1)."""
    s0, s1, ow, oh = C.c_size_t(), C.c_size_t(), C.c_size_t(), C.c_size_t()
    rc = lib().ipb_stripe_plan(C.byref(ops), C.byref(settings) if settings is not None else None, width, height,
                               out_row0, out_row1, C.byref(s0), C.byref(s1), C.byref(ow), C.byref(oh))
    _capi.check(None, rc)
    return s0.value, s1.value, ow.value, oh.value


def plan_stripes(ops, settings, width, height, world, multiple=2):
    """One StripeLayout per rank.  Source-row ownership is cut in the middle of each overlap, so both neighbours
    receive about half of the shared rows; halos never reach past the adjacent rank (checked)."""
    _, _, ow, oh = stripe_plan(ops, settings, width, height)
    parts = partition_rows(oh, world, multiple)
    need = []
    for r0, r1 in parts:
        if r1 > r0:
            s0, s1, _, _ = stripe_plan(ops, settings, width, height, r0, r1)
        else:
            s0 = s1 = need[-1][1] if need else 0
        need.append((s0, s1))
    own = [0]
    for k in range(1, world):
        lo, hi = need[k][0], need[k - 1][1]      # rows [lo, hi) are wanted by both stripes (may be empty)
        own.append(max(own[-1], min((lo + hi) // 2 if hi > lo else hi, height)))
    own.append(height)
    layouts = []
    for k, (r0, r1) in enumerate(parts):
        # a stripe's buffer always holds its own block; the kernel reads the sub-range it needs
        s0, s1 = own[k], own[k + 1]
        if r1 > r0:
            s0, s1 = min(s0, need[k][0]), max(s1, need[k][1])
        if (k > 0 and s0 < own[k - 1]) or (k + 1 < world and s1 > own[k + 2]):
            raise ValueError(f"stripe {k} needs source rows [{s0},{s1}) beyond its neighbours' blocks: too many ranks "
                             f"for a {height}-row frame")
        layouts.append(StripeLayout(k, r0, r1, s0, s1, own[k], own[k + 1], ow, oh))
    return layouts


class Comm:
    """ipb_comm: the NCCL communicator of the stripe ranks, created inside libipb200.so (which resolves libnccl at run
    time).  `unique_id()` on rank 0, hand the 128 bytes to every rank, then Comm(id, rank, nranks, device, stream)."""

    def __init__(self, uid, rank, nranks, device, stream):
        self.handle = C.c_void_p()
        rc = lib().ipb_comm_create(int(device), C.c_void_p(int(stream)) if stream else None, bytes(uid), int(rank),
                                   int(nranks), C.byref(self.handle))
        if rc != 0:
            raise _capi.IpbError(rc, (lib().ipb_comm_last_error(None) or b"").decode())
        self.rank, self.nranks = int(rank), int(nranks)

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(_capi.COMM_ID_BYTES)
        rc = lib().ipb_comm_unique_id(buf)
        if rc != 0:
            raise _capi.IpbError(rc, (lib().ipb_comm_last_error(None) or b"").decode())
        return buf.raw

    @staticmethod
    def nccl_version():
        v = C.c_int()
        return v.value if lib().ipb_comm_nccl_version(C.byref(v)) == 0 else None

    def close(self):
        if self.handle:
            lib().ipb_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def halo_plan(layouts, rank, row_bytes):
    """ipb_halo of stripe `rank`: which bytes of its buffer (rows src_row0.. of the frame) go to / come from the neighbours."""
    me = layouts[rank]
    h = _capi.Halo()
    for nb in (rank - 1, rank + 1):
        if nb < 0 or nb >= len(layouts):
            continue
        other = layouts[nb]
        sa, sb = (other.halo_down if nb < rank else other.halo_up)   # what the neighbour's stencil reads ...
        sa, sb = max(sa, me.own_row0), min(sb, me.own_row1)           # ... of the rows this rank owns
        ra, rb = (me.halo_up if nb < rank else me.halo_down)
        side = "up" if nb < rank else "down"
        if sb > sa:
            setattr(h, f"send_{side}_off", (sa - me.src_row0) * row_bytes)
            setattr(h, f"send_{side}_bytes", (sb - sa) * row_bytes)
        if rb > ra:
            setattr(h, f"recv_{side}_off", (ra - me.src_row0) * row_bytes)
            setattr(h, f"recv_{side}_bytes", (rb - ra) * row_bytes)
    return h


def exchange_halos_nccl(comm, ptrs, layouts, row_bytes):
    """The halo exchange through the C ABI (ipb_halo_exchange: grouped ncclSend / ncclRecv on the communicator's stream).
    `ptrs`: device addresses of the stripe buffers (one per frame; all share the layout).  No torch involved."""
    h = halo_plan(layouts, comm.rank, row_bytes)
    arr = (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])
    rc = lib().ipb_halo_exchange(comm.handle, arr, len(ptrs), C.byref(h))
    if rc != 0:
        raise _capi.IpbError(rc, (lib().ipb_comm_last_error(comm.handle) or b"").decode())


def exchange_halos(buf, layouts, rank, group=None):
    """Fill the halo rows of `buf` — a 2-D torch tensor (rows src_row0..src_row1 of the frame, any dtype, one row of
    samples per tensor row) whose own block is already in place — from the neighbouring ranks, and send them theirs.
    `buf` may also be a list of such tensors (several frames with the same layout): all their rows travel in ONE
    batched isend/irecv group, which amortises the NCCL launch latency over the frames (the messages are tiny:
    one or two rows).  Returns after the transfers are complete (on the GPU: enqueued on the current stream, which
    then waits for them — no host synchronisation)."""
    import torch
    import torch.distributed as dist
    me = layouts[rank]
    bufs = list(buf) if isinstance(buf, (list, tuple)) else [buf]
    ops = []
    for nb in (rank - 1, rank + 1):
        if nb < 0 or nb >= len(layouts):
            continue
        other = layouts[nb]
        # what the neighbour needs from my own block / what I need from the neighbour's own block
        sa, sb = (other.halo_down if nb < rank else other.halo_up)
        sa, sb = max(sa, me.own_row0), min(sb, me.own_row1)
        ra, rb = (me.halo_up if nb < rank else me.halo_down)
        for b in bufs:
            if b.element_size() != 1:  # NCCL has no 16-bit integer type: move the rows as bytes
                b = b.view(torch.uint8)
            if sb > sa:
                ops.append(dist.P2POp(dist.isend, b[sa - me.src_row0: sb - me.src_row0], nb, group))
            if rb > ra:
                ops.append(dist.P2POp(dist.irecv, b[ra - me.src_row0: rb - me.src_row0], nb, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return buf


class DevicePtr:
    """A caller-owned device allocation (e.g. a torch tensor's storage) usable as an output destination."""

    def __init__(self, ptr, nbytes, keep=None):
        self.ptr, self.nbytes, self._keep = int(ptr), int(nbytes), keep


def run_stripe_8bit(pipeline, rows, layout, dst):
    """Launch the fused kernel for `layout` on the pipeline's context.  `rows`: source rows [layout.src_row0,
    layout.src_row1) as u16 samples of the sensor width — a device address (int) or a host numpy array (then the
    banded H2D / kernel / D2H path runs); `dst`: DevicePtr / DeviceArray / numpy array for the stripe's 8-bit rows."""
    from .pipeline import ImageSource
    img = pipeline.globals.image
    data = rows if hasattr(rows, "ctypes") else int(rows)
    src = ImageSource(_capi.SRC_RAW_U16, img.width, layout.src_row1 - layout.src_row0, 1, data)
    pipeline.set_stripe_source(src, layout.src_row0, layout.out_row0, layout.out_row1)
    return pipeline.output_8bit_stripe(dst=dst, rows=layout.out_row1 - layout.out_row0, width=layout.out_width)


def run_stripes_8bit_batch(pipeline, rows_ptr, nframes, layout, dst):
    """The stripes `layout` of nframes frames, stacked in one device buffer (frame k's source rows [layout.src_row0,
    layout.src_row1) start k * (src_row1 - src_row0) rows after device address rows_ptr), in one call of
    ipb_pipeline_output_8bit_batch; the results follow each other in dst (DevicePtr / DeviceArray)."""
    from .pipeline import ImageSource
    img = pipeline.globals.image
    rows = layout.src_row1 - layout.src_row0
    src = ImageSource(_capi.SRC_RAW_U16, img.width, rows, 1, int(rows_ptr))
    pipeline.set_stripe_source(src, layout.src_row0, layout.out_row0, layout.out_row1)
    return pipeline.output_8bit_batch(nframes, rows, dst, (layout.out_row1 - layout.out_row0) * layout.out_width * 3)
