"""CPU checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/tci_b200.h declares, and refuses to work (no CPU fallback) when no device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def T():
    import __graft_entry__ as g
    g.build()
    import tci_b200
    return tci_b200


def header_symbols():
    src = open(os.path.join(ROOT, "include", "tci_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tci_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(T):
    names = header_symbols()
    assert len(names) >= 25
    L = ctypes.CDLL(os.path.join(ROOT, "tensorcrossinterpolation.jl_b200", "libtci_b200.so"))
    for n in names:
        assert hasattr(L, n), n
    from tci_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names  # the binding covers exactly the header


def test_version_and_no_cpu_fallback(T):
    import torch
    assert T.lib().tci_version() >= 100
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TCIError, match="no CPU fallback"):
        T.Context(0)
    with pytest.raises(T.TCIError):
        T.BuiltinTarget(T.LORENTZ, [1.0], [10, 10])


def test_product_does_not_import_oracle():
    """The product path must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "tensorcrossinterpolation.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for fn in files:
            if not fn.endswith((".py", ".cu", ".h")):
                continue
            for line in open(os.path.join(dirpath, fn)):
                assert "libtci_oracle" not in line and "orc_" not in line, (fn, line)
                if "oracle" in line:
                    assert "import" not in line and "include" not in line and "CDLL" not in line, (fn, line)
