#!/usr/bin/env python
"""bench.py — megapixels/s of the raw->sRGB hot path (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --steps K --warmup W    the reference's CPU path (oracle port) on host cores

Workload (config.workload): BASELINE config 2 — 6000x4000 RGGB Bayer frames, full pipe to 8-bit sRGB
(gofloat -> demosaic -> to_lab -> basecurve -> from_lab -> gamma -> pack).  A step is one pass of the hot path
over a batch of FRAMES_PER_STEP frames that rotate through NSETS distinct input/output buffer sets
(NSETS * 120 MB > the 126 MB L2, so no frame is served from cache).  At N > 1 every rank runs the same batch on
its own GPU (frames are independent: no data-path collective, scaling "weak"); value = total MP/s.

  --workload c3 | c4   BASELINE configurations 3 (8256x5504 X-Trans) and 4 (24 MP frames with the 4x down-scale)
                       through the same harness
  --workload c5        BASELINE configuration 5: one 11648x8736 frame cut into row stripes, one per rank, NCCL halo
                       exchange of the stencil rows + one fused launch per rank per frame (scaling "strong")

One JSON line on stdout (rank 0): value (device-resident, CUDA events on the launching stream, max over ranks),
roofline (algorithmic bytes / measured launch time against MEASURED_PEAKS.json), e2e (pinned host buffers through
the synchronous API, copies inside the timed region), cpu_baseline (the oracle's port of the reference CPU path on
the host cores, N = 1 only), gpu_launches, clocks.  See DESIGN.md "Measurement" for the arithmetic.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H = 6000, 4000
MP = W * H / 1e6
CFA = "RGGB"
SETTINGS = {}          # PipelineSettings overrides of the workload (c4: maxwidth / maxheight)
OUT_W, OUT_H = W, H    # size of the result
WORKLOAD_NAME = "C2: 6000x4000 RGGB Bayer -> 8-bit sRGB, fused demosaic->gamma kernel"
KERNEL_NAME, TRAFFIC_KEY = "k_spec8<512> (speculative 8-bit kernel)", "k_spec8 C2 6000x4000"
NSETS = 8
FRAMES_PER_STEP = 32
E2E_THREADS = 2
ALGO_BYTES_PER_PX = 5  # 2 B in (u16 CFA sample) + 3 B out (u8 sRGB) — SURVEY.md §8d


def select_workload(name):
    """c2 is the contract's line; c3 / c4 are the other single-GPU BASELINE configurations, run the same way."""
    global W, H, MP, CFA, SETTINGS, OUT_W, OUT_H, WORKLOAD_NAME, KERNEL_NAME, TRAFFIC_KEY, ALGO_BYTES_PER_PX
    if name == "c3":
        import common
        W, H, CFA = 8256, 5504, common.XTRANS
        OUT_W, OUT_H = W, H
        WORKLOAD_NAME = "C3: 8256x5504 Fuji X-Trans 6x6 -> 8-bit sRGB, fused demosaic->gamma kernel (generic CFA path)"
        KERNEL_NAME, TRAFFIC_KEY = "k_spec8<512, generic CFA> (speculative 8-bit kernel)", "k_spec8 C3 8256x5504"
    elif name == "c4":
        SETTINGS = {"maxwidth": 1500, "maxheight": 1000}
        OUT_W, OUT_H = 1500, 1000
        WORKLOAD_NAME = ("C4: 6000x4000 RGGB Bayer -> 1500x1000 8-bit sRGB (4x down-scale inside the demosaic op, "
                         "scaled_demosaic), frames round-robin over the GPUs")
        KERNEL_NAME, TRAFFIC_KEY = "k_spec8_scaled (speculative 8-bit kernel behind scaled_demosaic)", "k_spec8_scaled C4 6000x4000->1500x1000"
        ALGO_BYTES_PER_PX = 2 + 3.0 * OUT_W * OUT_H / (W * H)  # per INPUT pixel: 2.1875
    MP = W * H / 1e6

METRIC = "megapixels/sec raw->sRGB full pipe"


def workload_config(world, frames_per_step):
    """config of the JSON line — the same dict on both arms (the reference arm times bounded samples of this workload)."""
    return {"workload": WORKLOAD_NAME, "frames_per_step": frames_per_step, "frames_per_step_all_gpus": world * frames_per_step,
            "buffer_sets": NSETS,
            "l2": f"inputs larger than L2: {NSETS} rotating sets x {(W * H * 2 + OUT_W * OUT_H * 3) / 1e6:.0f} MB",
            "parallelism": f"frames round-robin, {world} replica(s), no collective"}


def workload_params():
    import common
    return common.raw_params(cfa=CFA)


def measured_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[key]
        return e["dram_read_bytes"] + e["dram_write_bytes"]
    except Exception:
        return None


def measured_issue(key):
    """{warp_inst_per_px, issue_active} of the dominant kernel from the same committed capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            e = json.load(f)[key]
        return {"warp_inst_per_px": e["warp_inst_per_px"], "issue_active": e["issue_active"], "source": e.get("capture")}
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "applications_clocks_setting": 0x2, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_to_gpu_numa(index):
    """Pin this process (and the pinned host buffers it allocates afterwards: first touch) to the CPUs of the NUMA node
    the GPU hangs off, when the box tells (sysfs).  Returns the node, or None when unknown / single-node."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def pinned_array(ip, nbytes, dtype, shape):
    p = C.c_void_p()
    rc = ip.lib().ipb_host_alloc(nbytes, C.byref(p))
    if rc != 0:
        raise RuntimeError("ipb_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


def cpu_reference_leg(steps, warmup):
    """Times the oracle's op-by-op CPU pipeline (same pass structure as the reference: one pass and one allocation per
    op, clones before basecurve / from_lab / gamma, serial pack loop) on all host cores, one frame of the workload per
    step, built on this machine with BASELINE.md's flags (-O3 -march=native -ffp-contract=off)."""
    import common
    import oracle
    oracle.build()
    data = common.synth_cfa(W, H)
    small = common.synth_cfa(640, 360)
    want_small = oracle.pipeline_output_8bit(oracle.make_pipeline(small, "raw", workload_params(), SETTINGS or None))
    flags = oracle.use_native()
    L = oracle.lib()
    L.orc_set_threads(0)
    cores = L.orc_get_threads()
    same = bool(np.array_equal(want_small, oracle.pipeline_output_8bit(
        oracle.make_pipeline(small, "raw", workload_params(), SETTINGS or None))))
    p = oracle.make_pipeline(data, "raw", workload_params(), SETTINGS or None)
    for _ in range(warmup):
        oracle.pipeline_output_8bit(p)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        oracle.pipeline_output_8bit(p)
        times.append(time.perf_counter() - t0)
    per = float(np.mean(times))
    return {"value": MP / per, "unit": "MP/s", "cores": int(cores), "kind": "port",
            "sample": f"{steps} steps x one {W}x{H} frame of the workload through output_8bit; C port of the reference CPU "
                      f"path (Rust toolchain absent), OpenMP over rows on {int(cores)} threads, gcc {flags}; "
                      f"same bytes as the checker build on a 640x360 frame: {same}",
            "ms_per_frame": per * 1e3, "flags": flags}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (its C port) on this box's host cores, one frame of the workload per
    step, the driver's --steps / --warmup, the same metric and config as our arm.  Rank 0 alone works."""
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cb = cpu_reference_leg(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "MP/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_frame"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(max(world, args.gpus), args.frames_per_step),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def strong_leg(args, ip, common, torch, dist, ctx, stream, barrier, rank, world, local_rank, params, K, Wm):
    """BASELINE config 5 (strong scaling): one 11648x8736 RGGB frame cut into row stripes, one per rank.  The stencil
    rows travel through the C ABI's halo exchange (ipb_halo_exchange: grouped ncclSend / ncclRecv between stripe
    neighbours on the compute stream; torch.distributed only carries the 128-byte NCCL id to the ranks), then every
    rank runs one fused launch on its stripe.  Returns the dict of the JSON line's "strong" key (rank 0) — with
    `parity`: every rank's stripe equals, byte for byte, the same rows of a single-launch run of the whole frame."""
    from imagepipe_b200 import _capi
    from imagepipe_b200.sharded import (Comm, DevicePtr, exchange_halos_nccl, halo_plan, plan_stripes, run_stripe_8bit,
                                        run_stripes_8bit_batch)
    W5, H5 = 11648, 8736
    mp5 = W5 * H5 / 1e6
    dummy = ip.DeviceArray(64, ctx)
    p = ip.Pipeline.new_from_source(ip.ImageSource(_capi.SRC_RAW_U16, W5, H5, 1, dummy), ctx=ctx)
    common.fill_ipb_ops(p.ops, params)
    lays = plan_stripes(p.ops, p.globals.settings, W5, H5, world)
    me = lays[rank]
    rows_out = me.out_row1 - me.out_row0
    set_bytes = (me.src_row1 - me.src_row0) * W5 * 2 + rows_out * W5 * 3
    nsets = max(2, -(-600_000_000 // set_bytes))  # rotating working set of >= 600 MB per GPU (L2 is 126 MB)
    comm = None
    if world > 1:
        box = [Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = Comm(box[0], rank, world, local_rank, stream.cuda_stream)
    src_rows_me = me.src_row1 - me.src_row0
    with torch.cuda.stream(stream):
        # the stripes of the nsets frames in flight sit one after the other in one allocation (so do their results): they
        # can be converted by one batched launch after one grouped exchange
        in_block = torch.zeros((nsets, src_rows_me, W5), dtype=torch.int16, device="cuda")
        out_block = torch.empty((nsets, rows_out, W5, 3), dtype=torch.uint8, device="cuda")
        bufs = [in_block[i] for i in range(nsets)]
        outs = [out_block[i] for i in range(nsets)]
        for i in range(nsets):
            own = bufs[i][me.own_row0 - me.src_row0: me.own_row1 - me.src_row0]
            ip.lib().ipb_synth_cfa_u16(ctx.handle, common.SEED + i, W5, me.own_row0, me.own_row1 - me.own_row0, own.data_ptr())
    ptrs = [b.data_ptr() for b in bufs]
    # the stripes of a step's frames in one batched launch (profiles/r02e_batch.txt: 72.3 against 76.9 us per stripe at
    # 8 GPUs, no difference at 2); --no-batch: one launch per stripe
    batched = not args.no_batch
    hp = halo_plan(lays, rank, W5 * 2)
    halo_bytes = hp.send_up_bytes + hp.send_down_bytes + hp.recv_up_bytes + hp.recv_down_bytes

    # A step is F = nsets frames, each in its own buffer set.  Their halo rows (one raw row per neighbour and frame)
    # travel in one NCCL group at the head of the step, then every frame is one fused launch: the exchange latency
    # (tens of microseconds, against < 0.1 ms of compute per stripe at 8 GPUs) is paid once per step.
    F = nsets

    out_all = DevicePtr(out_block.data_ptr(), out_block.numel())

    def step():
        if comm is not None:
            exchange_halos_nccl(comm, ptrs, lays, W5 * 2)
        if batched:   # the F stripes in one launch (ipb_pipeline_output_8bit_batch on the stripe source)
            run_stripes_8bit_batch(p, ptrs[0], F, me, out_all)
        else:
            for j in range(F):
                run_stripe_8bit(p, ptrs[j], me, DevicePtr(outs[j].data_ptr(), outs[j].numel()))

    with torch.cuda.stream(stream):
        for _ in range(Wm):
            step()
    barrier()
    # One step is captured into a CUDA graph and replayed: at 8 GPUs a stripe takes < 0.1 ms and the host path per
    # frame (Python + ctypes, ~0.2 ms) would otherwise be the bottleneck.
    graph, mode = None, "eager"
    if not args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                step()
            graph, mode = g, "cuda graph of one step"
        except Exception as e:  # noqa: BLE001
            print(f"bench.py: graph capture failed on rank {rank} ({e}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()
    if world > 1:
        modes = [None] * world
        dist.all_gather_object(modes, mode)
        if any(m != modes[0] for m in modes) or modes[0] == "eager":
            graph, mode = None, "eager"   # all ranks must agree, or the exchange would not pair up

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            step()

    g = None
    with torch.cuda.stream(stream):
        run_step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(K):
            run_step()
        ev[1].record(stream)
    barrier()
    t = torch.tensor([ev[0].elapsed_time(ev[1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = K * F * mp5 / (total_ms / 1e3)

    # parity: the whole frame of buffer set 0 in one launch on this GPU, against this rank's stripe of the sharded run
    with torch.cuda.stream(stream):
        full = ip.synth_cfa_u16(common.SEED, W5, 0, H5, ctx=ctx)
        pf = ip.Pipeline.new_from_source(ip.ImageSource.Raw(full, width=W5, height=H5, cpp=1), ctx=ctx)
        common.fill_ipb_ops(pf.ops, params)
        ref = torch.empty((H5, W5, 3), dtype=torch.uint8, device="cuda")
        pf.output_8bit(dst=DevicePtr(ref.data_ptr(), ref.numel()))
        ok = int(torch.equal(ref[me.out_row0:me.out_row1], outs[0])) if rows_out else 1
    torch.cuda.synchronize()
    okt = torch.tensor([int(ok)], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    del ref, full

    # e2e: every rank's source rows (halo included: the host holds the whole frame) in pinned host memory, its output
    # stripe back to pinned host memory, through the banded H2D / kernel / D2H path
    src_rows = me.src_row1 - me.src_row0
    host_in, _hin = pinned_array(ip, src_rows * W5 * 2, np.uint16, (src_rows, W5))
    host_out, _hout = pinned_array(ip, rows_out * W5 * 3, np.uint8, (rows_out, W5, 3))
    with torch.cuda.stream(stream):
        if comm is not None:
            exchange_halos_nccl(comm, ptrs[:1], lays, W5 * 2)
    torch.cuda.synchronize()
    host_in[:] = bufs[0].cpu().numpy().view(np.uint16)
    e2e_calls = 6
    for _ in range(2):
        run_stripe_8bit(p, host_in, me, host_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        run_stripe_8bit(p, host_in, me, host_out)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_calls * mp5 / float(te.item())
    same = bool(np.array_equal(host_out, outs[0].cpu().numpy()))
    # the captured step holds the communicator: drop the graph before the communicator goes
    graph = None
    torch.cuda.synchronize()
    if comm is not None:
        barrier()
        comm.close()
    launch_ms = total_ms / (K * F)
    return {
        "workload": "C5: 11648x8736 RGGB Bayer -> 8-bit sRGB, one frame cut into row stripes (one per GPU), halo rows "
                    "by ipb_halo_exchange (NCCL send/recv), " +
                    ("the stripes of the step's frames in one batched launch" if batched else "one fused launch per stripe"),
        "value": value, "unit": "MP/s", "scaling": "strong", "steps": K, "frames_per_step": F,
        "ms_per_step": total_ms / K, "ms_per_frame": launch_ms, "stripe_rows": rows_out,
        "halo_rows": (me.own_row0 - me.src_row0) + (me.src_row1 - me.own_row1), "halo_bytes": int(halo_bytes),
        "launch": mode, "launches_per_step": 1 if batched else F, "nccl": Comm.nccl_version() if world > 1 else None,
        "buffer_sets": nsets, "set_mb": set_bytes / 1e6,
        "achieved_gbs": ALGO_BYTES_PER_PX * rows_out * W5 / (launch_ms / 1e3) / 1e9,
        "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_frame": src_rows * W5 * 2,
                "d2h_bytes_per_frame": rows_out * W5 * 3, "frames_timed": e2e_calls, "matches_device_path": same,
                "note": "bytes are per rank; every rank copies its own stripe"},
        "parity": bool(int(okt.item()) == 1),
        "parity_what": "each rank's stripe == the same rows of a single-launch run of the whole frame on that rank's GPU",
    }


def run_c5(args, ip, common, torch, dist, ctx, stream, barrier, rank, world, local_rank, params, K, Wm, F):
    """--workload c5: the strong-scaling leg as the line's headline."""
    sampler = ClockSampler(local_rank)
    sampler.start()
    st = strong_leg(args, ip, common, torch, dist, ctx, stream, barrier, rank, world, local_rank, params, K, Wm)
    clocks = sampler.stop()
    if rank == 0:
        peak, peak_kind = measured_peak()
        line = {
            "metric": METRIC, "value": st["value"], "unit": "MP/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": st["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": st["workload"], "frames_per_step": st["frames_per_step"], "buffer_sets": st["buffer_sets"],
                       "stripe_rows": st["stripe_rows"], "halo_rows": st["halo_rows"],
                       "l2": f"inputs larger than L2: {st['buffer_sets']} rotating sets x {st['set_mb']:.0f} MB per GPU",
                       "launch": st["launch"], "parallelism": f"{world} row stripe(s)"},
            "roofline": {"bound": "hbm", "achieved": st["achieved_gbs"], "peak": peak, "unit": "GB/s",
                         "frac": st["achieved_gbs"] / peak, "traffic": None,
                         "kernel": "k_spec8 (one stripe, exchange included in the time)", "kernel_ms": st["ms_per_frame"],
                         "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * st["stripe_rows"] * 11648},
            "e2e": {"value": st["e2e"]["value"], "unit": "MP/s", "h2d_bytes_per_step": st["e2e"]["h2d_bytes_per_frame"],
                    "d2h_bytes_per_step": st["e2e"]["d2h_bytes_per_frame"], "matches_device_path": st["e2e"]["matches_device_path"]},
            "gpu_launches": int(K * st["launches_per_step"]), "clocks": clocks, "parity": st["parity"],
        }
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--frames-per-step", type=int, default=None)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling (C5 stripes) leg of the default line")
    ap.add_argument("--no-batch", action="store_true",
                    help="one output_8bit call (one launch) per frame instead of ipb_pipeline_output_8bit_batch over the buffer sets")
    ap.add_argument("--no-graph", action="store_true", help="c5: launch every frame from Python instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"],
                    help="c2 (default, the contract's line): 24 MP frames, replicas; c3: 45 MP X-Trans frames; c4: 24 MP frames "
                         "with the 4x down-scale; c5: one 101.8 MP frame per step-frame, row stripes over the ranks with an "
                         "NCCL halo exchange (strong scaling)")
    args = ap.parse_args()
    # a hung collective must not hang the caller: dump every thread's stack and exit after IPB_BENCH_WATCHDOG seconds
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("IPB_BENCH_WATCHDOG", "1500")), exit=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    select_workload(args.workload)
    if args.frames_per_step is None:
        args.frames_per_step = 8 if args.workload == "c5" else FRAMES_PER_STEP
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import common
    import imagepipe_b200 as ip

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()
    ctx = ip.Context(local_rank, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, Wm, F = args.steps, max(3, args.warmup), args.frames_per_step
    params = workload_params()
    if args.workload == "c5":
        run_c5(args, ip, common, torch, dist, ctx, stream, barrier, rank, world, local_rank, params, K, Wm, F)
        if world > 1:
            dist.destroy_process_group()
        return
    # NSETS distinct synthetic frames, generated on the device, and NSETS output buffers, each kind in
    # one allocation so that the sets can also be handed over as one batch
    from imagepipe_b200.sharded import DevicePtr
    frame_bytes, out_bytes = W * H * 2, OUT_W * OUT_H * 3
    frames_block = ip.DeviceArray(NSETS * frame_bytes, ctx)
    outs_block = ip.DeviceArray(NSETS * out_bytes, ctx)
    for i in range(NSETS):
        ip.lib().ipb_synth_cfa_u16(ctx.handle, common.SEED + rank * 1000 + i, W, 0, H, frames_block.ptr + i * frame_bytes)
    outs = [DevicePtr(outs_block.ptr + i * out_bytes, out_bytes, keep=outs_block) for i in range(NSETS)]
    pipes = []
    for i in range(NSETS):
        src = ip.ImageSource(ip._capi.SRC_RAW_U16, W, H, 1, frames_block.ptr + i * frame_bytes, keep=frames_block)
        p = ip.Pipeline.new_from_source(src, ctx=ctx)
        common.fill_ipb_ops(p.ops, params)
        for k, v in SETTINGS.items():
            setattr(p.globals.settings, k, v)
        assert p.output_size() == (OUT_W, OUT_H)
        pipes.append(p)

    # A step is F frames.  Per frame: one Pipeline.output_8bit call (the reference's API shape), one launch.  Batched
    # (default): ipb_pipeline_output_8bit_batch over the NSETS frames at a time — one launch per batch where the
    # full-resolution speculative kernel applies (start-up and tail once per batch), the same per-frame launches inside
    # the C call elsewhere (C4).
    batched = not args.no_batch and F % NSETS == 0

    def step_frames():
        for j in range(F):
            pipes[j % NSETS].output_8bit(dst=outs[j % NSETS])

    def step_batches():
        for _ in range(F // NSETS):
            pipes[0].output_8bit_batch(NSETS, H, outs_block, out_bytes)

    step = step_batches if batched else step_frames

    for _ in range(Wm):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    ctx.spec_stats(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for k in range(K):
            step()
            ev[k + 1].record(stream)
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    st = ctx.spec_stats()
    # pixels recomputed exactly per launch, as a fraction of the pixels a launch produces (every workload runs a
    # speculative kernel now: k_spec8 on C2 / C3, k_spec8_scaled on C4)
    spec_stats = {"recomputed_fraction": st["fixups"] / max(1, K * F) / (OUT_W * OUT_H), "certified_delta": st["delta"],
                  "xu_cbrt_rel_err": st["mufu_err"]}
    total_ms = ev[0].elapsed_time(ev[K])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(K)]
    per_frame_calls = None
    if batched:   # the same frames through one output_8bit call each (device resident), for comparison
        kk = max(2, min(K, 5))
        step_frames()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.cuda.stream(stream):
            pe[0].record(stream)
            for _ in range(kk):
                step_frames()
            pe[1].record(stream)
        pe[1].synchronize()
        pms = pe[0].elapsed_time(pe[1]) / (kk * F)
        per_frame_calls = {"ms_per_frame": pms, "value_per_gpu": MP / (pms / 1e3), "unit": "MP/s",
                           "what": "one Pipeline.output_8bit call (one launch) per frame, device resident, this rank"}
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * K * F * MP / (total_ms_max / 1e3)

    # ---- e2e: the same metric through Pipeline.output_8bit (synchronous, like the reference's) with HOST buffers
    # (pinned), H2D and D2H inside the timed region.  E2E_THREADS host threads, each with its own context, pipeline and
    # pinned buffers, call it in a loop — the reference lets several Pipelines run concurrently on different threads
    #  — so that one frame's D2H overlaps the next frame's H2D; whole-frame copies (band_mb 0) use the
    # PCIe link best in that regime (tools/pcie_dep_probe.py).  The latency of a single banded call is reported too.
    e2e_frames, e2e_steps = 4, max(5, K // 5)
    frame0 = frames_block.to_numpy(np.uint16)[: W * H].reshape(H, W)
    workers = []
    for t in range(E2E_THREADS):
        wctx = ctx if t == 0 else ip.Context(local_rank)
        host_in, _hp_in = pinned_array(ip, W * H * 2, np.uint16, (H, W))
        host_out, _hp_out = pinned_array(ip, OUT_W * OUT_H * 3, np.uint8, (OUT_H, OUT_W, 3))
        host_in[:] = frame0
        pe = ip.Pipeline.new_from_source(ip.ImageSource.Raw(host_in), ctx=wctx)
        common.fill_ipb_ops(pe.ops, params)
        for k, v in SETTINGS.items():
            setattr(pe.globals.settings, k, v)
        workers.append((wctx, pe, host_in, host_out))
    pe, host_out = workers[0][1], workers[0][3]
    for _ in range(2):
        pe.output_8bit(dst=host_out)
    t0 = time.perf_counter()
    for _ in range(8):
        pe.output_8bit(dst=host_out)  # banded H2D + kernel + D2H, synchronous at return
    single_call_ms = (time.perf_counter() - t0) / 8 * 1e3
    for w in workers:
        w[1].set_band_mb(0)
        w[1].output_8bit(dst=w[3])
    barrier()
    per_thread = e2e_steps * e2e_frames // E2E_THREADS
    go = threading.Barrier(E2E_THREADS + 1)

    def e2e_loop(pw, out):
        go.wait()
        for _ in range(per_thread):
            pw.output_8bit(dst=out)

    ths = [threading.Thread(target=e2e_loop, args=(w[1], w[3])) for w in workers]
    for th in ths:
        th.start()
    go.wait()
    t0 = time.perf_counter()
    for th in ths:
        th.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * per_thread * E2E_THREADS * MP / float(te.item())
    # result check on the last e2e frame: the device-resident path produced the same bytes
    same = all(bool(np.array_equal(w[3], outs_block.to_numpy(np.uint8)[:out_bytes].reshape(OUT_H, OUT_W, 3))) for w in workers)
    # the box's copy ceiling for this traffic at N GPUs: every rank moves one frame's bytes (48 MB in, 72 MB out, pinned,
    # both directions at once on the two copy streams) with nothing else, all ranks at once
    ceil_in = torch.empty(W * H * 2, dtype=torch.uint8, device="cuda")
    ceil_out = torch.empty(OUT_W * OUT_H * 3, dtype=torch.uint8, device="cuda")
    hin_t, hout_t = torch.from_numpy(workers[0][2].reshape(-1).view(np.uint8)), torch.from_numpy(workers[0][3].reshape(-1))
    cs1, cs2 = torch.cuda.Stream(), torch.cuda.Stream()

    def raw_copies():
        with torch.cuda.stream(cs1):
            ceil_in.copy_(hin_t, non_blocking=True)
        with torch.cuda.stream(cs2):
            hout_t.copy_(ceil_out, non_blocking=True)

    for _ in range(4):
        raw_copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(40):
        raw_copies()
    torch.cuda.synchronize()
    tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    pcie_ceiling = world * 40 * MP / float(tc.item())

    # ---- strong scaling on the same record: BASELINE config 5, one 101.8 MP frame over the N GPUs (every N, N = 1 too)
    strong = None
    if args.workload == "c2" and not args.no_strong:
        del workers
        strong = strong_leg(args, ip, common, torch, dist, ctx, stream, barrier, rank, world, local_rank, params,
                            max(20, min(K, 50)), 5)   # at 8 GPUs a step is < 1 ms: enough of them for a stable figure

    if rank == 0:
        peak, peak_kind = measured_peak()
        # the step contains nothing but its launches: launches / K of them, each over frames_per_launch frames
        per_step = max(1, int(launches) // K)
        frames_per_launch = F / per_step
        kernel_ms = float(np.mean(step_ms)) / per_step
        achieved = ALGO_BYTES_PER_PX * W * H * frames_per_launch / (kernel_ms / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, F),   # the same dict on the reference arm
            "launches": ("ipb_pipeline_output_8bit_batch over %d frames at a time" % NSETS) if batched
            else "one Pipeline.output_8bit call per frame",
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (measured_traffic(TRAFFIC_KEY) or 0) * frames_per_launch or None, "kernel": KERNEL_NAME,
                         "kernel_ms": kernel_ms, "peak_kind": peak_kind,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX * W * H * frames_per_launch,
                         "frames_per_launch": frames_per_launch, "kernel_ms_per_frame": kernel_ms / frames_per_launch,
                         "issue": measured_issue(TRAFFIC_KEY), "spec": spec_stats},
            "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": e2e_frames * W * H * 2,
                    "d2h_bytes_per_step": e2e_frames * OUT_W * OUT_H * 3, "steps": e2e_steps, "frames_per_step": e2e_frames,
                    "host_threads": E2E_THREADS, "frames_timed": per_thread * E2E_THREADS,
                    "single_call_ms": single_call_ms, "matches_device_path": same,
                    "numa_node": numa_node, "pcie_ceiling": pcie_ceiling, "frac_of_ceiling": e2e_value / pcie_ceiling,
                    "pcie_ceiling_what": "the same bytes per frame as bare pinned-memory copies, both directions at once, "
                                         "all ranks at the same time (no kernel, no API)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if per_frame_calls is not None:
            line["per_frame_calls"] = per_frame_calls
        if strong is not None:
            line["strong"] = strong
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference_leg(3, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
