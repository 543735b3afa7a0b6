"""The C-ABI library loads on a CPU-only machine and exports every symbol include/laghos_b200.h
declares (no compute calls without a GPU); the product fails loudly without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "laghos_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lagb_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    from laghos_b200._lib import LIB_PATH, SYMBOLS
    lib = ctypes.CDLL(LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/laghos_b200.h but not exported"
    assert sorted(SYMBOLS) == syms, set(SYMBOLS) ^ set(syms)
    for s in ("lagb_laghos_run", "lagb_run_options_default"):
        assert hasattr(lib, s)


def test_host_setup_without_gpu(built):
    """Host-side setup works anywhere; device entry points refuse to run without CUDA."""
    import torch
    from laghos_b200.api import Problem, Context, LagbError
    P = Problem("cube01_hex", 1, 1, 2, 1)
    assert P.NE == 64 and P.ndofs_h1 == 9 ** 3 and P.ndofs_l2 == 64 * 8
    if not torch.cuda.is_available():
        with pytest.raises(LagbError):
            Context(P)


def test_product_does_not_import_oracle():
    """The product path must never route through the oracle (prompt, section 3)."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "laghos_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"pyoracle|laghos_oracle|oracle/|liboracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
