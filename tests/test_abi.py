"""CPU test: the C-ABI library loads and exports every symbol include/dbg_b200.h declares (no compute
calls without a GPU), and the product path fails loudly without a device instead of falling back."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dbg_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    from rust_debruijn_b200 import _lib
    L = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/dbg_b200.h but not exported by libdbg_b200.so"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in rust_debruijn_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback():
    import torch

    import rust_debruijn_b200 as D
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(D.DbgError) as e:
        D.Context(0)
    assert e.value.status == 3  # DBG_E_CUDA


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "rust_debruijn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f


def test_integration_md_lists_every_entry_point():
    """INTEGRATION.md's Rust extern block is generated from the header (tools/gen_rust_extern.py) and must stay in step with it;
    dbg_stats / dbg_multi_info are declared field for field (a zero-sized placeholder would let dbg_stats_get write past it)."""
    import subprocess
    import sys
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for s in header_symbols():
        assert f"pub fn {s}(" in doc, f"{s} missing from INTEGRATION.md"
    gen = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "gen_rust_extern.py")]).decode()
    assert gen in doc, "INTEGRATION.md extern block is stale: regenerate with tools/gen_rust_extern.py"
    assert "pub struct dbg_stats { pub n_seqs: u64" in doc and "pub struct dbg_multi_info { pub n_ranks: u32" in doc
