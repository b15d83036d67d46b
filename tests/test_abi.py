"""The C-ABI library loads and exports every symbol include/elimrec_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from elimrec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "elimrec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(elimrec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m elimrec_b200.build`"
    l = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(l, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_version_and_error_channel():
    l = _lib.lib()
    assert l.elimrec_abi_version() == 2
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = l.elimrec_spmm(100, 0, 1, 0, None, None, None, None, None, None, 4, None, 4, None, None, None)
    assert rc == -1 and b"width" in l.elimrec_last_error()
    rc = l.elimrec_rank_topk(None, 1, None, None, None, None, 99, None, None, None)
    assert rc == -1


def test_no_cpu_path():
    import pytest
    import torch
    from elimrec_b200._lib import ElimrecError, ptr
    with pytest.raises(ElimrecError):
        ptr(torch.zeros(4))
