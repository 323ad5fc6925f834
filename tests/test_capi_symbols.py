"""The C-ABI library loads on a box without a GPU and exports every function include/ipb200.h declares; the
ctypes binding declares a signature for each; the host-only entry points (size negotiation, stripe planning)
work without a device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "ipb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = header_functions()
    assert len(names) >= 60
    for must in ("ipb_ctx_create", "ipb_gofloat_run", "ipb_demosaic_run", "ipb_rotatecrop_run", "ipb_tolab_run",
                 "ipb_basecurve_run", "ipb_fromlab_run", "ipb_gamma_run", "ipb_transform_run", "ipb_pipeline_run",
                 "ipb_pipeline_output_8bit", "ipb_pipeline_output_16bit", "ipb_stripe_plan"):
        assert must in names


def test_library_exports_every_declared_symbol(ip):
    L = ip.lib()
    missing = [n for n in header_functions() if not hasattr(L, n)]
    assert not missing, f"libipb200.so does not export {missing}"


def test_binding_declares_every_symbol(ip):
    sigs = ip.lib()._ipb_signatures
    missing = [n for n in header_functions() if n not in sigs]
    assert not missing, f"_capi.py has no signature for {missing}"


def test_struct_sizes_match_header(ip):
    from imagepipe_b200 import _capi
    # the sizes the C compiler gives the PODs of ipb200.h on LP64 (size_t = 8, natural alignment)
    assert C.sizeof(_capi.GoFloat) == 4 * 8 + 4 + 16 + 16 + 4
    assert C.sizeof(_capi.Demosaic) == 148
    assert C.sizeof(_capi.ToLab) == (12 + 12 + 12 + 4) * 4
    assert C.sizeof(_capi.BaseCurve) == 8 + 8 + 32 * 2 * 4
    assert C.sizeof(_capi.Settings) == 4 * 8 + 8
    assert C.sizeof(_capi.Stripe) == 32


def test_host_only_entry_points_need_no_gpu(ip):
    assert ip.lib().ipb_version() == 100
    assert ip.scaling_size(6000, 4000, 1500, 1000) == (1500, 1000)
    assert ip.scaling_size(100, 100, 0, 0) == (100, 100)
    assert ip.calculate_scale(6000, 4000, 1500, 1000) == pytest.approx(4.0)


def test_no_cpu_fallback_without_a_device(ip):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ip.IpbError) as e:
        ip.Context(0)
    assert "CUDA" in str(e.value)


def test_header_is_plain_c_and_cxx(tmp_path):
    """include/ipb200.h compiles as C11 and as C++17 on its own (no CUDA / torch types in the ABI)."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "ipb200.h")
    src = tmp_path / "t.c"
    src.write_text('#include "ipb200.h"\nint main(void) { ipb_settings s = {0}; ipb_ops o; (void)o; return (int)s.maxwidth + (IPB_VERSION == 0); }\n')
    for cmd in (["gcc", "-std=c11", "-Wall", "-Werror", "-pedantic", "-fsyntax-only"], ["g++", "-x", "c++", "-std=c++17", "-Wall", "-fsyntax-only"]):
        r = subprocess.run(cmd + ["-I", os.path.dirname(hdr), str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
