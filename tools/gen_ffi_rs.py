#!/usr/bin/env python
"""Generates shim-rs/src/ffi.rs — the Rust `extern "C"` block and `#[repr(C)]` twins — from include/ipb200.h.

  python tools/gen_ffi_rs.py            writes shim-rs/src/ffi.rs
  python tools/gen_ffi_rs.py --check    exits 1 when the committed file differs from what the header gives

The header is the single source of truth of the boundary; tests/test_shim_rs.py runs the check and also compares
symbols, field names and field order of the two files independently of this generator's output format."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "ipb200.h")
OUT = os.path.join(ROOT, "shim-rs", "src", "ffi.rs")

SCALAR = {"size_t": "usize", "int": "c_int", "float": "f32", "double": "f64", "char": "c_char", "uint8_t": "u8",
          "uint16_t": "u16", "uint32_t": "u32", "uint64_t": "u64", "unsigned long long": "c_ulonglong",
          "unsigned char": "u8", "long": "c_long", "void": "c_void"}


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def rust_type(ctype):
    """C type (without declarator name) -> Rust type."""
    t = " ".join(ctype.split())
    const = False
    ptrs = t.count("*")
    t = t.replace("*", " ").strip()
    parts = t.split()
    # `T *const *` (pointer to const pointer) -> treat the inner const as part of the pointee chain
    if "const" in parts:
        const = True
        parts = [p for p in parts if p != "const"]
    base = " ".join(p for p in parts if p != "struct")
    r = SCALAR.get(base, base)
    for _ in range(ptrs):
        r = ("*const " if const else "*mut ") + r
    return r


def parse(src):
    src = strip_comments(src)
    structs, enums, funcs, defines, opaque = [], [], [], [], []
    for m in re.finditer(r"#define\s+(IPB_[A-Z0-9_]+)\s+(\d+)", src):
        defines.append((m.group(1), int(m.group(2))))
    for m in re.finditer(r"typedef struct (\w+) (\w+);", src):
        opaque.append(m.group(2))
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} (\w+);", src, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            tm = re.match(r"((?:const )?(?:unsigned )?(?:long long|\w+)(?: \*)*) ?(.*)", decl)
            ctype, names = tm.group(1), tm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                stars = nm.count("*")
                nm = nm.replace("*", "").strip()
                dims = [int(defines_dict(defines).get(d, d)) for d in re.findall(r"\[(\w+)\]", nm)]
                nm = re.sub(r"\[.*", "", nm)
                rt = rust_type(ctype + " *" * stars)
                for d in reversed(dims):
                    rt = f"[{rt}; {d}]"
                fields.append((nm, rt))
        structs.append((m.group(3), fields))
    for m in re.finditer(r"(?:typedef )?enum (?:\w+ )?\{(.*?)\}", src, flags=re.S):
        for item in m.group(1).split(","):
            item = " ".join(item.split())
            if "=" in item:
                k, v = item.split("=")
                enums.append((k.strip(), int(v.strip())))
    body = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
    body = re.sub(r"(?:typedef )?enum (?:\w+ )?\{.*?\}[^;]*;", "", body, flags=re.S)
    for m in re.finditer(r"^((?:const )?(?:unsigned )?(?:long long|\w+)(?: \*)?)\s*\*?\s*(ipb_\w+)\(([^;{}]*?)\);", body, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if "*" in m.group(0).split(name)[0] and "*" not in ret:
            ret += " *"
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                am = re.match(r"(.*?)(\w+)((?:\[\w*\])*)$", a)
                ctype, pname, arr = am.group(1).strip(), am.group(2), am.group(3)
                if arr:
                    ctype += " *"
                params.append((pname, rust_type(ctype)))
        funcs.append((name, params, None if ret == "void" else rust_type(ret)))
    return defines, opaque, structs, enums, funcs


def defines_dict(defines):
    return {k: str(v) for k, v in defines}


def render():
    defines, opaque, structs, enums, funcs = parse(open(HDR).read())
    done = {s for s, _ in structs}
    out = ["// ffi.rs — GENERATED from include/ipb200.h by tools/gen_ffi_rs.py (do not edit; tests/test_shim_rs.py checks it).",
           "// The `extern \"C\"` block the reference (pedrocr/imagepipe) binds libipb200.so with: one declaration per C",
           "// entry point, `#[repr(C)]` twins of the parameter structs (which mirror the reference's serde structs field",
           "// by field: see the comments of ipb200.h), the status and enum constants.  NOT compiled here: this image has",
           "// no Rust toolchain (INTEGRATION.md).",
           "#![allow(non_camel_case_types, dead_code)]",
           "use std::os::raw::{c_char, c_int, c_long, c_ulonglong, c_void};", ""]
    for k, v in defines:
        out.append(f"pub const {k}: usize = {v};")
    for k, v in enums:
        out.append(f"pub const {k}: c_int = {v};")
    out.append("")
    for o in opaque:
        if o not in done:
            out.append(f"#[repr(C)] pub struct {o} {{ _private: [u8; 0] }}")
    out.append("")
    for name, fields in structs:
        out.append("#[repr(C)]\n#[derive(Clone, Copy)]")
        out.append(f"pub struct {name} {{")
        for f, t in fields:
            out.append(f"    pub {f}: {t},")
        out.append("}\n")
    out.append('#[link(name = "ipb200")]\nextern "C" {')
    for name, params, ret in funcs:
        ps = ", ".join(f"{('r#' + p) if p in ('in', 'type', 'ref', 'box', 'move') else p}: {t}" for p, t in params)
        out.append(f"    pub fn {name}({ps})" + (f" -> {ret};" if ret else ";"))
    out.append("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = render()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "w").write(text)
    print(f"wrote {OUT}: {text.count('pub fn ')} functions, {text.count('pub struct ')} structs")
