"""Generate the Rust `extern "C"` block of INTEGRATION.md from include/dbg_b200.h: one declaration per entry point, and the
two plain-data structs (dbg_stats, dbg_multi_info) field for field.  python tools/gen_rust_extern.py > /tmp/extern.rs"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "include", "dbg_b200.h")).read()
src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

SCALAR = {"int": "c_int", "uint64_t": "u64", "uint32_t": "u32", "uint16_t": "u16", "uint8_t": "u8", "int64_t": "i64", "float": "f32",
          "char": "c_char", "void": "c_void"}


def rust_type(ct):
    ct = ct.strip()
    const = False
    stars = ct.count("*")
    base = ct.replace("*", " ")
    toks = [t for t in base.split() if t not in ("const", "struct")]
    const = "const" in base.split()
    name = toks[0]
    rt = SCALAR.get(name, name)
    if stars == 0:
        return rt
    out = rt
    # innermost pointer constness follows the leading const; outer levels are mutable (out parameters), except `const T* const*`
    parts = ct.split("*")
    for i in range(stars):
        seg_const = "const" in parts[i].split() if i == 0 else "const" in parts[i].split()
        out = ("*const " if seg_const else "*mut ") + out
    return out


def structs():
    out = []
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(dbg_stats|dbg_multi_info)\s*;", src, flags=re.S):
        body, name = m.group(1), m.group(2)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            for n in names.split(","):
                fields.append(f"pub {n.strip()}: {SCALAR[ty]}")
        out.append(f"#[repr(C)] #[derive(Clone, Copy, Default)] pub struct {name} {{ " + ", ".join(fields) + " }")
    return out


def functions():
    out = []
    for m in re.finditer(r"^\s*((?:const\s+)?[a-z_0-9]+\s*\*?)\s*(dbg_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)([A-Za-z_][A-Za-z_0-9]*)$", a)
                ty, pn = mm.group(1), mm.group(2)
                if pn == "in":
                    pn = "input"
                params.append(f"{pn}: {rust_type(ty)}")
        r = "" if ret == "void" else " -> " + rust_type(ret)
        out.append(f"    pub fn {name}({', '.join(params)}){r};")
    return out


if __name__ == "__main__":
    fns = functions()
    print("// debruijn-b200-sys/src/lib.rs — generated from include/dbg_b200.h by tools/gen_rust_extern.py")
    print("use std::os::raw::{c_char, c_int, c_void};")
    for h in ("dbg_ctx", "dbg_seqset", "dbg_kmer_table", "dbg_graph", "dbg_partition", "dbg_comm", "dbg_multi"):
        print(f"#[repr(C)] pub struct {h} {{ _p: [u8; 0] }}")
    for s_ in structs():
        print(s_)
    print(f"\n// {len(fns)} entry points")
    print('#[link(name = "dbg_b200")]\nextern "C" {')
    print("\n".join(fns))
    print("}")
