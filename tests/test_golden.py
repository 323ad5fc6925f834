"""Golden vectors of tests/golden/ (an independent numpy-f32 restatement of the reference, see make_golden.py) against
the CPU oracle (CPU suite) and against the CUDA path through the C ABI (GPU suite).

Every comparison is bit-exact (0 ulp; the north star allows 1e-5 relative): gofloat / demosaic::full /
scaled_demosaic are pure f32 arithmetic, and the colour chain's tables are built from the same glibc cbrtf / powf in
the golden generator, the oracle and the product's host code.
"""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import common
from common import assert_bit_exact

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEMOSAIC_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz"))
                        if not os.path.basename(p).startswith("colour"))
def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: d[k] for k in d.files}


def params_of(g):
    return common.raw_params(cfa=str(g["cfa"]), crops=tuple(int(v) for v in g["crops"]), black=float(g["black"]),
                             white=float(g["white"]))


def close(got, want, what):
    assert_bit_exact(np.asarray(got, np.float32), np.asarray(want, np.float32), what)


def test_fixtures_are_present():
    assert len(DEMOSAIC_CASES) == 5 and os.path.exists(os.path.join(GOLD, "colour_chain_96px.npz"))


def test_worked_example_by_hand():
    """full_rggb_10x10 pixel (1,1) is blue in an RGGB mosaic: R = mean of the 4 diagonal reds, G = mean of the 4 edge
    greens, B = the sample itself; pixel (0,0) is a red corner: G = mean of its 2 in-frame greens, B = the one blue."""
    g = load("full_rggb_10x10")
    a, d = g["gofloat"], g["demosaic"]
    f = np.float32
    assert d[1, 1, 2] == a[1, 1]
    assert d[1, 1, 0] == f(f(f(f(a[0, 0] + a[0, 2]) + a[2, 0]) + a[2, 2]) / f(4))
    assert d[1, 1, 1] == f(f(f(f(a[0, 1] + a[1, 0]) + a[1, 2]) + a[2, 1]) / f(4))
    assert d[0, 0, 0] == a[0, 0] and d[0, 0, 1] == f(f(a[0, 1] + a[1, 0]) / f(2)) and d[0, 0, 2] == a[1, 1]
    assert d[..., 3].max() == 0  # no fourth colour in RGGB: the E channel stays 0
    raw = g["raw"]
    assert a[0, 0] == f(f(f(raw[0, 0]) - f(64)) / f(1023 - 64)) and a[0, 0] < 0   # below black: negative survives
    assert a[2, 3] == 1.0                                                        # at white


# ------------------------------------------------------------------------------------------------ oracle (CPU)

@pytest.mark.parametrize("name", DEMOSAIC_CASES)
def test_oracle_matches_golden_demosaic(orc, name):
    g = load(name)
    ops = orc.Ops()
    orc.fill_ops(ops, params_of(g))
    src, keep = orc.make_source(np.ascontiguousarray(g["raw"]), "raw", 1)
    gf = orc.lib().orc_gofloat_run(C.byref(ops.gofloat), C.byref(src))
    got_gf, _ = orc.buffer_to_numpy(gf, free=False)
    assert_bit_exact(got_gf[..., 0], g["gofloat"], f"{name}: oracle gofloat")
    cfa = orc.Cfa()
    assert orc.lib().orc_cfa_new(C.byref(cfa), str(g["cfa"]).encode()) == 0
    if "nwidth" in g:
        dm = orc.lib().orc_scaled_demosaic(C.byref(cfa), gf, int(g["nwidth"]), int(g["nheight"]))
    else:
        dm = orc.lib().orc_demosaic_full(C.byref(cfa), gf)
    orc.lib().orc_buffer_free(gf)
    assert_bit_exact(orc.buffer_to_numpy(dm)[0], g["demosaic"], f"{name}: oracle demosaic")


def test_oracle_matches_golden_colour_chain(orc):
    g = load("colour_chain_96px")
    ops = orc.Ops()
    orc.fill_ops(ops, common.raw_params(matrix=g["matrix"], wb=[float(v) for v in g["wb"]],
                                        points=[tuple(p) for p in g["points"]]))
    buf = orc.buffer_from_numpy(g["rgbe"].reshape(8, 12, 4))
    lab = orc.lib().orc_tolab_run(C.byref(ops.tolab), buf)
    cur = orc.lib().orc_basecurve_run(C.byref(ops.basecurve), lab)
    close(orc.buffer_to_numpy(cur, free=False)[0].reshape(96, 3), g["lab"], "oracle lab after basecurve")
    rgb = orc.lib().orc_fromlab_run(cur)
    st = orc.Settings()
    out = orc.lib().orc_gamma_run(C.byref(st), rgb)
    close(orc.buffer_to_numpy(out, free=False)[0].reshape(96, 3), g["rgb"], "oracle rgb after gamma")


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)

@pytest.mark.gpu
@pytest.mark.parametrize("name", DEMOSAIC_CASES)
def test_cuda_matches_golden_demosaic(ip, ctx, name):
    g = load(name)
    st = {"maxwidth": int(g["nwidth"]), "maxheight": int(g["nheight"])} if "nwidth" in g else {}
    p = common.make_ipb_pipeline(ip, np.ascontiguousarray(g["raw"]), "raw", params_of(g), st, ctx=ctx)
    if st:
        p.output_size()  # the size walk sets settings.demosaic_width/height (pipeline.rs:331-338)
    gf = p.ops.gofloat.run(p.globals)
    assert_bit_exact(gf.to_numpy()[..., 0], g["gofloat"], f"{name}: cuda gofloat")
    dm = p.ops.demosaic.run(p.globals, gf)
    assert_bit_exact(dm.to_numpy(), g["demosaic"], f"{name}: cuda demosaic")


@pytest.mark.gpu
def test_cuda_matches_golden_colour_chain(ip, ctx):
    g = load("colour_chain_96px")
    p = common.make_ipb_pipeline(ip, np.zeros((16, 16), np.uint16), "raw",
                                 common.raw_params(matrix=g["matrix"], wb=[float(v) for v in g["wb"]],
                                                   points=[tuple(p) for p in g["points"]]), ctx=ctx)
    buf = ip.OpBuffer.from_numpy(g["rgbe"].reshape(8, 12, 4), ctx=ctx)
    cur = p.ops.basecurve.run(p.globals, p.ops.tolab.run(p.globals, buf))
    close(cur.to_numpy().reshape(96, 3), g["lab"], "cuda lab after basecurve")
    out = p.ops.gamma.run(p.globals, p.ops.fromlab.run(p.globals, cur))
    close(out.to_numpy().reshape(96, 3), g["rgb"], "cuda rgb after gamma")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["full_xtrans_12x12", "full_gbrg_11x13", "scaled_xtrans_30x24_to_10x8"])
def test_cuda_fused_matches_golden_through_the_chain(ip, orc, ctx, name):
    """The fused kernel from the golden raw frame: its result equals the colour chain applied op by op to the golden
    demosaic output (bit-exact: same tables on both sides)."""
    g = load(name)
    st = {"maxwidth": int(g["nwidth"]), "maxheight": int(g["nheight"])} if "nwidth" in g else {}
    p = common.make_ipb_pipeline(ip, np.ascontiguousarray(g["raw"]), "raw", params_of(g), st, ctx=ctx)
    fused = p.run().to_numpy()
    buf = ip.OpBuffer.from_numpy(g["demosaic"], ctx=ctx)
    cur = p.ops.basecurve.run(p.globals, p.ops.tolab.run(p.globals, buf))
    want = p.ops.gamma.run(p.globals, p.ops.fromlab.run(p.globals, cur)).to_numpy()
    assert_bit_exact(fused, want, f"{name}: fused vs golden demosaic + per-op chain")


def test_committed_fixtures_are_what_the_generator_writes(tmp_path, monkeypatch):
    """tests/golden/*.npz are exactly the output of tests/golden/make_golden.py (the independent numpy restatement)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "HERE", str(tmp_path))
    mod.main()
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "*.npz")))
    assert names == sorted(os.listdir(tmp_path))
    for n in names:
        a, b = np.load(os.path.join(GOLD, n)), np.load(os.path.join(tmp_path, n))
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            if a[k].dtype.kind == "f":
                assert np.array_equal(a[k].view(np.uint32) if a[k].dtype == np.float32 else a[k], b[k].view(np.uint32) if b[k].dtype == np.float32 else b[k], equal_nan=False) or np.array_equal(a[k], b[k], equal_nan=True), (n, k)
            else:
                assert np.array_equal(a[k], b[k]), (n, k)
