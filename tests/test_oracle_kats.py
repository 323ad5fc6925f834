"""The reference's own known-answer tests, run against the CPU oracle (oracle/).  These pin the restatement:
every assert here is an assert of the reference's test-suite, cited by file:line (paths under the reference)."""
import ctypes as C

import numpy as np
import pytest

F_ROWS = ["        ", " RRRRRR ", " GG     ", " BBBB   ", " GG     ", " GG     ", "        "]  # transform.rs:153-165
# transform.rs:167-278: golden outputs per rawloader Orientation
ORIENT_GOLD = {
    "Normal": F_ROWS, "Unknown": F_ROWS,
    "HorizontalFlip": ["        ", " RRRRRR ", "     GG ", "   BBBB ", "     GG ", "     GG ", "        "],
    "VerticalFlip": ["        ", " GG     ", " GG     ", " BBBB   ", " GG     ", " RRRRRR ", "        "],
    "Rotate90": ["       ", " GGBGR ", " GGBGR ", "   B R ", "   B R ", "     R ", "     R ", "       "],
    "Rotate270": ["       ", " R     ", " R     ", " R B   ", " R B   ", " RGBGG ", " RGBGG ", "       "],
    "Rotate180": ["        ", "     GG ", "     GG ", "   BBBB ", "     GG ", " RRRRRR ", "        "],
    "Transpose": ["       ", " RGBGG ", " RGBGG ", " R B   ", " R B   ", " R     ", " R     ", "       "],
    "Transverse": ["       ", "     R ", "     R ", "   B R ", "   B R ", " GGBGR ", " GGBGR ", "       "],
}
# rawloader Orientation::to_flips -> (transpose, flip_x, flip_y); derived from the golden bitmaps above
TO_FLIPS = {"Normal": (0, 0, 0), "Unknown": (0, 0, 0), "VerticalFlip": (0, 0, 1), "HorizontalFlip": (0, 1, 0),
            "Rotate180": (0, 1, 1), "Transpose": (1, 0, 0), "Rotate90": (1, 0, 1), "Rotate270": (1, 1, 0),
            "Transverse": (1, 1, 1)}
# OpTransform::new (transform.rs:24-36): Orientation -> (rotation, fliph, flipv), and the flips OpTransform::run
# (transform.rs:56-66) derives from those fields.  Note the reference's own quirk: a Transverse file becomes
# (Rotate270, fliph) whose flips are (t, t^1, f) == Transpose; the restatement follows run(), not the intent.
OP_FIELDS = {"Normal": (0, 0, 0), "Unknown": (0, 0, 0), "VerticalFlip": (0, 0, 1), "HorizontalFlip": (0, 1, 0),
             "Rotate180": (2, 0, 0), "Transpose": (1, 0, 1), "Rotate90": (1, 0, 0), "Rotate270": (3, 0, 0),
             "Transverse": (3, 1, 0)}
RUN_FLIPS = dict(TO_FLIPS, Transverse=(1, 0, 0))


def rgb_str(rows):
    t = {"R": (1, 0, 0), "G": (0, 1, 0), "B": (0, 0, 1), "O": (1, 1, 1), " ": (0, 0, 0)}
    return np.array([[t[c] for c in r] for r in rows], np.float32)


@pytest.mark.parametrize("kat", ["roundtrip_8bit", "roundtrip_16bit", "roundtrip_8bit_gamma", "roundtrip_16bit_gamma",
                                 "roundtrip_8bit_lab_xyz", "roundtrip_8bit_lab_rgb", "roundtrip_8bit_lab_rgb_gamma",
                                 "roundtrip_16bit_lab_xyz", "roundtrip_16bit_lab_rgb", "roundtrip_16bit_lab_rgb_gamma",
                                 "rotatecrop_roundtrip_transform", "rotatecrop_roundtrip_transform_rotation"])
def test_reference_roundtrip_kats(orc, kat):
    """color_conversions.rs:337-349,385-402,420-611 and rotatecrop.rs:273-312 — zero mismatches."""
    assert getattr(orc.lib(), "orc_kat_" + kat)() == 0


def test_curves_kats(orc):
    """curves.rs:164-189"""
    L = orc.lib()

    def spline(pts):
        s = orc.Spline()
        arr = np.array(pts, np.float32).reshape(-1, 2)
        L.orc_spline_new(C.byref(s), arr.ctypes.data, len(pts))
        return lambda v: L.orc_spline_interpolate(C.byref(s), v)
    f = spline([])
    assert f(0.0) == 0.0 and f(1.0) == 1.0          # extremes
    assert f(1.5) == 1.0 and f(-0.2) == 0.0         # saturates
    assert spline([(0.0, 0.2)])(0.0) == np.float32(0.2)   # high_blackpoint
    assert spline([(1.0, 0.8)])(1.0) == np.float32(0.8)   # low_whitepoint


def test_default_basecurve_coefficients(orc):
    """f32 coefficients of the raw default curve [(0.5, 0.6)]"""
    s = orc.Spline()
    arr = np.array([(0.5, 0.6)], np.float32)
    orc.lib().orc_spline_new(C.byref(s), arr.ctypes.data, 1)
    hexs = lambda a, n: [np.float32(a[i]).view(np.uint32).item() for i in range(n)]
    assert hexs(s.c1, 3) == [0x3f99999a, 0x3f75c28f, 0x3f4ccccc]
    assert hexs(s.c2, 2) == [0x3ef5c290, 0xbf23d70e]
    assert hexs(s.c3, 2) == [0xbf75c290, 0x3f23d710]


@pytest.mark.parametrize("name", sorted(ORIENT_GOLD))
def test_transform_orientation_kats(orc, name):
    """transform.rs:167-278 via rotate_buffer(flips) and via OpTransform's (rotation, fliph, flipv) fields."""
    L = orc.lib()
    src = rgb_str(F_ROWS)
    want = rgb_str(ORIENT_GOLD[name])
    b = orc.buffer_from_numpy(src)
    got, _ = orc.buffer_to_numpy(L.orc_rotate_buffer(b, *TO_FLIPS[name]))
    assert np.array_equal(got, want)
    op = orc.Transform(*OP_FIELDS[name])
    flips = (C.c_int * 3)()
    L.orc_orientation_flips(C.byref(op), flips)
    assert tuple(flips) == RUN_FLIPS[name]
    L.orc_buffer_free(b)


def test_scaling_noop(orc):
    """scaling.rs:188-203"""
    w = h = 150
    data = np.arange(w * h * 3, dtype=np.uint32).astype(np.uint16)
    out = np.empty_like(data)
    orc.lib().orc_scale_down_srgb16(data.ctypes.data, w, h, w, h, out.ctypes.data)
    assert np.array_equal(out, data)


def rc_setup(orc):
    a = np.arange(100 * 100 * 3, dtype=np.float32).reshape(100, 100, 3)  # rotatecrop.rs:175-183
    return a, orc.buffer_from_numpy(a)


@pytest.mark.parametrize("crops,size,first", [
    (dict(crop_top=0.1), (100, 90), 100 * 10 * 3), (dict(crop_bottom=0.1), (100, 90), 0),
    (dict(crop_top=0.1, crop_bottom=0.1), (100, 80), 100 * 10 * 3), (dict(crop_left=0.1), (90, 100), 10 * 3),
    (dict(crop_right=0.1), (90, 100), 0), (dict(crop_left=0.1, crop_right=0.1), (80, 100), 10 * 3),
    (dict(crop_left=0.1, crop_right=0.1, crop_top=0.1, crop_bottom=0.1), (80, 80), 100 * 10 * 3 + 10 * 3),
    (dict(rotation=0.5), (141, 141), None), (dict(rotation=1.0), (100, 100), None)])
def test_rotatecrop_kats(orc, crops, size, first):
    """rotatecrop.rs:185-271"""
    a, b = rc_setup(orc)
    op = orc.RotateCrop(0, 0, 0, 0, 0, 1.0, 0, 0, 0)
    for k, v in crops.items():
        setattr(op, k, v)
    got, _ = orc.buffer_to_numpy(orc.lib().orc_rotatecrop_run(C.byref(op), b))
    assert (got.shape[1], got.shape[0]) == size
    if first is not None:
        assert got.reshape(-1)[0] == a.reshape(-1)[first]
    orc.lib().orc_buffer_free(b)


def all_colors_8bit():
    v = np.arange(256, dtype=np.uint8)
    r, g, b = np.meshgrid(v, v, v, indexing="ij")
    return np.stack([r, g, b], -1).reshape(4096, 4096, 3)


@pytest.mark.parametrize("fast", [True, False])
def test_roundtrip_8bit_all_colors(orc, fast):
    """tests/roundtrip_test.rs:4-35 — every (R,G,B) u8 through the whole pipeline comes back unchanged."""
    img = all_colors_8bit()
    p = orc.make_pipeline(img, "rgb", settings={"use_fastpath": fast})
    out = orc.pipeline_output_8bit(p)
    assert np.array_equal(out, img)


def blocks_16bit():
    """tests/roundtrip_test.rs:37-72: strided (89, 97, 101) u16 colours in 4096x4096 blocks."""
    r = np.arange(0, 65536, 89, dtype=np.uint16)
    g = np.arange(0, 65536, 97, dtype=np.uint16)
    b = np.arange(0, 65536, 101, dtype=np.uint16)
    return r, g, b


def block_16bit(idx):
    """Block `idx` of the reference's enumeration, laid out exactly like its loop fills image_data (the reference
    restarts g and b from the break position; every colour triple of the strided grid is visited)."""
    r, g, b = blocks_16bit()
    per = 4096 * 4096
    total = r.size * g.size * b.size
    lo, hi = idx * per, min((idx + 1) * per, total)
    lin = np.arange(lo, hi, dtype=np.int64)
    out = np.zeros((per, 3), np.uint16)
    out[: hi - lo, 0] = r[lin // (g.size * b.size)]
    out[: hi - lo, 1] = g[(lin // b.size) % g.size]
    out[: hi - lo, 2] = b[lin % b.size]
    return out.reshape(4096, 4096, 3), -(-total // per)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("blk", [0, -1])
def test_roundtrip_16bit_blocks(orc, fast, blk):
    """tests/roundtrip_test.rs:37-84 (first and last block here; the GPU suite runs every block)."""
    _, nblocks = block_16bit(0)
    img, _ = block_16bit(blk % nblocks)
    p = orc.make_pipeline(img, "rgb", settings={"use_fastpath": fast})
    out = orc.pipeline_output_16bit(p)
    assert np.array_equal(out, img)


MAXSIZE_CASES = [  # tests/maxsize_test.rs:32-90 (source 128x64)
    ({}, {}, (128, 64)),
    ({"maxwidth": 128}, {}, (128, 64)),
    ({"maxwidth": 64}, {}, (64, 32)),
    ({"maxwidth": 64}, {"transform": {"rotation": 1}}, (64, 128)),
    ({"maxwidth": 32}, {"transform": {"rotation": 1}}, (32, 64)),
    ({"maxwidth": 256}, {"transform": {"rotation": 1}}, (64, 128)),
    ({"maxwidth": 64}, {"gofloat": {"crop_top": 1, "crop_bottom": 1, "crop_left": 1, "crop_right": 1}}, (64, 31)),
    ({"maxwidth": 64}, {"rotatecrop": {"crop_top": 0.1, "crop_bottom": 0.1, "crop_left": 0.1, "crop_right": 0.1}},
     (64, 32)),
]


@pytest.mark.parametrize("settings,params,size", MAXSIZE_CASES)
def test_maxsize_kats(orc, settings, params, size):
    img = np.zeros((64, 128, 3), np.uint8)
    for fast in (True, False):
        p = orc.make_pipeline(img, "rgb", params, dict(settings, use_fastpath=fast))
        out8 = orc.pipeline_output_8bit(p)
        assert (out8.shape[1], out8.shape[0]) == size
        p = orc.make_pipeline(img, "rgb", params, dict(settings, use_fastpath=fast))
        out16 = orc.pipeline_output_16bit(p)
        assert (out16.shape[1], out16.shape[0]) == size
