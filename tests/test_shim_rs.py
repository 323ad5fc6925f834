"""shim-rs/ (the Rust side of the boundary, uncompiled: no toolchain in this image) stays in step with include/ipb200.h:
same symbols, same struct fields in the same order, and the hand-written gpu.rs only uses what exists."""
import ctypes as C
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "ipb200.h")
FFI = os.path.join(ROOT, "shim-rs", "src", "ffi.rs")
GPU = os.path.join(ROOT, "shim-rs", "src", "gpu.rs")


def _hdr():
    return re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)


def header_functions():
    return re.findall(r"\b(ipb_\w+)\s*\(", re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", _hdr(), flags=re.S))


def header_structs():
    out = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \w+;", _hdr(), flags=re.S):
        names = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            rest = re.sub(r"^(?:const )?(?:unsigned )?(?:long long|\w+)\s*", "", " ".join(decl.split()), count=1)
            for nm in rest.split(","):
                names.append(re.sub(r"[\*\s]|\[.*", "", nm))
        out[m.group(1)] = names
    return out


def rust_functions():
    return re.findall(r"pub fn (ipb_\w+)\(", open(FFI).read())


def rust_structs():
    out = {}
    text = open(FFI).read()
    for m in re.finditer(r"pub struct (\w+) \{\n(.*?)\n\}", text, flags=re.S):
        out[m.group(1)] = re.findall(r"pub (\w+):", m.group(2))
    for m in re.finditer(r"pub struct (\w+) \{ _private: \[u8; 0\] \}", text):
        out[m.group(1)] = ["_private"]
    return out


def test_ffi_rs_is_what_the_header_generates():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_ffi_rs.py"), "--check"])
    assert r.returncode == 0, "shim-rs/src/ffi.rs is stale: run python tools/gen_ffi_rs.py"


def test_symbols_and_order_match_the_header():
    h, r = header_functions(), rust_functions()
    assert len(h) > 80 and h == r


def test_struct_fields_and_order_match_the_header():
    h, r = header_structs(), rust_structs()
    assert set(h) <= set(r)
    for name, fields in h.items():
        assert r[name] == fields, name
    # the opaque handles exist as zero-sized #[repr(C)] types
    for opaque in ("ipb_ctx", "ipb_buffer", "ipb_pipeline", "ipb_cache", "ipb_comm"):
        assert r[opaque] == ["_private"]


def test_gpu_rs_uses_only_what_ffi_rs_declares():
    src = open(GPU).read()
    funcs, structs = set(rust_functions()), rust_structs()
    used = set(re.findall(r"\b(ipb_\w+)\b", src))
    known = funcs | set(structs)
    assert used <= known, sorted(used - known)
    # all eight ops implement the trait, with the reference's op names (src/ops/*.rs name())
    impls = re.findall(r"impl<'a, 'c> ImageOp<'a> for Gpu<'c, (\w+)>", src)
    assert impls == ["OpGoFloat", "OpDemosaic", "OpRotateCrop", "OpToLab", "OpBaseCurve", "OpFromLab", "OpGamma", "OpTransform"]
    assert re.findall(r'fn name\(&self\) -> &str \{ "(\w+)" \}', src) == ["gofloat", "demosaic", "rotatecrop", "to_lab", "basecurve",
                                                                         "from_lab", "gamma", "transform"]
    for fn in ("transform_forward", "transform_reverse", "reset", "run_cached", "output_8bit_cached", "output_16bit_cached"):
        assert fn in src
    # struct literals name every field of the twin they build
    for name, fields in structs.items():
        for m in re.finditer(r"(?<!-> )\b" + name + r" \{(.*?)\}", src, flags=re.S):
            body = m.group(1)
            if ":" not in body and "," not in body and len(fields) > 1:
                continue
            words = set(re.findall(r"\b(\w+)\b", body))   # `field: value` and the `field` shorthand alike
            assert set(fields) <= words, (name, sorted(set(fields) - words))


def test_library_exports_every_symbol_of_the_extern_block():
    lib = C.CDLL(os.path.join(ROOT, "imagepipe_b200", "libipb200.so"))
    for fn in rust_functions():
        assert hasattr(lib, fn), fn
