"""Known answers for the Lanczos extension's CPU statement (oracle/lanczos.c).  The reference has no Lanczos resampler
(scaling.rs:101-103 is a FIXME), so there is nothing of the reference's to pin this against: the definition itself is
checked — weights, their normalisation and symmetry, constants, impulse response, identity."""
import ctypes as C

import numpy as np
import pytest


def weights(orc, n_in, n_out, a=3):
    ks = C.c_size_t()
    orc.lib().orc_lanczos_weights(n_in, n_out, a, None, None, None, C.byref(ks))
    start, count = (C.c_int * n_out)(), (C.c_int * n_out)()
    w = np.zeros(n_out * ks.value, np.float32)
    orc.lib().orc_lanczos_weights(n_in, n_out, a, start, count, w.ctypes.data_as(C.POINTER(C.c_float)), None)
    return np.array(start), np.array(count), w.reshape(n_out, ks.value)


def resize(orc, arr, nw, nh, a=3):
    bp = orc.buffer_from_numpy(arr)
    out = orc.lib().orc_lanczos_resize(bp, nw, nh, a)
    orc.lib().orc_buffer_free(bp)
    assert out
    return orc.buffer_to_numpy(out)[0]


@pytest.mark.parametrize("n_in,n_out,a", [(6000, 1500, 3), (1000, 333, 3), (100, 150, 3), (64, 64, 2), (4000, 1000, 4)])
def test_weights_definition(orc, n_in, n_out, a):
    start, count, w = weights(orc, n_in, n_out, a)
    scale = n_in / n_out
    fs = max(scale, 1.0)
    assert w.shape[1] == int(np.ceil(a * fs)) * 2 + 1
    assert (count >= 1).all() and (start >= 0).all() and (start + count <= n_in).all()
    assert np.all(np.diff(start) >= 0)
    np.testing.assert_allclose(w.sum(1), 1.0, atol=3e-7)            # normalised
    i = n_out // 2                                                   # an interior sample against the closed form
    centre = (i + 0.5) * scale
    t = (np.arange(start[i], start[i] + count[i]) - centre + 0.5) / fs
    ref = np.sinc(t) * np.sinc(t / a) * (np.abs(t) < a)
    np.testing.assert_allclose(w[i, :count[i]], ref / ref.sum(), atol=1e-7)
    if n_in % n_out == 0 or n_in == n_out:                           # integer scale: the taps are symmetric
        inside = w[i, :count[i]][np.abs(t) < a - 1e-9]               # the window may carry one zero tap at |t| = a
        np.testing.assert_allclose(inside, inside[::-1], atol=1e-7)


def test_same_size_is_identity(orc):
    rng = np.random.default_rng(1)
    a = rng.random((37, 53, 3), dtype=np.float32)
    np.testing.assert_allclose(resize(orc, a, 53, 37), a, atol=2e-7)   # L(0) = 1, L(+-1), L(+-2) = O(1e-17)


def test_constant_image_stays_constant(orc):
    a = np.full((60, 90, 4), 0.3125, np.float32)
    for nw, nh in ((30, 20), (45, 17), (120, 77)):
        out = resize(orc, a, nw, nh)
        assert out.shape == (nh, nw, 4)
        np.testing.assert_allclose(out, 0.3125, rtol=0, atol=1e-6)


def test_impulse_response_is_the_tap_table(orc):
    """An impulse at column c of a single row: output column i holds w[i][c - start[i]]."""
    n_in, n_out = 200, 50
    start, count, w = weights(orc, n_in, n_out)
    a = np.zeros((1, n_in, 1), np.float32)
    a[0, 101, 0] = 1.0
    out = resize(orc, a, n_out, 1)[0, :, 0]
    for i in range(n_out):
        k = 101 - start[i]
        want = w[i, k] if 0 <= k < count[i] else 0.0
        assert out[i] == np.float32(want)


def test_downscale_removes_the_nyquist_pattern(orc):
    """A one-pixel checkerboard (the highest frequency) averages to its mean when reduced 4x — the low-pass property
    the reference's FIXME asks for."""
    y, x = np.mgrid[0:128, 0:128]
    a = ((x + y) & 1).astype(np.float32)[:, :, None]
    out = resize(orc, a, 32, 32)
    assert np.abs(out[4:-4, 4:-4] - 0.5).max() < 0.02
