"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/optex_b200.h declares, and refuses to compute without a B200 (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "optex_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(optex_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from optimaltextures_b200 import build

    return build.build()


def test_header_declares_what_the_binding_binds():
    from optimaltextures_b200 import _lib

    assert header_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_header_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (optex_[a-z0-9_]+)", out))
    missing = [s for s in header_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    handle = ctypes.CDLL(built_lib)
    for s in header_symbols():
        assert getattr(handle, s) is not None
    assert handle.optex_abi_version() == 1


def test_workspace_queries_need_no_gpu(built_lib):
    from optimaltextures_b200 import _lib

    lib = _lib.lib()
    n, c = 128 * 128, 512
    assert lib.optex_ot_workspace_bytes(n, n, c, _lib.MODES["cdf"]) >= 2 * 4 * n * c
    assert lib.optex_ot_workspace_bytes(n, n, c, 99) == 0          # unknown mode
    assert lib.optex_cdf_match_workspace_bytes(c, 256) >= 4 * 2 * c * 257
    assert lib.optex_rotation_workspace_bytes(c) >= 8 * (c - 1) * c


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_fallback_without_gpu(built_lib):
    import optimaltextures_b200 as ob
    from optimaltextures_b200 import _lib

    assert _lib.lib().optex_device_check() == _lib.EDEVICE
    assert _lib.lib().optex_last_error()
    x = torch.zeros(1, 4, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.optimal_transport(x, x, "cdf", rotation=torch.eye(8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ob.hist_match(x, x, "chol")
    # the raw C call reports EDEVICE instead of touching memory
    rc = _lib.lib().optex_cdf_match(None, None, None, 1, 1, 1, 256, None, None, 0, None)
    assert rc == _lib.EDEVICE


def test_mode_strings_follow_the_reference():
    from optimaltextures_b200 import _lib

    assert [_lib.mode_id(m) for m in ("chol", "pca", "sym", "cdf", "sort")] == [0, 1, 2, 3, 4]
    assert _lib.mode_id("anything-else") == _lib.MODES["sym"]   # histmatch.py:36 `else:  # sym`
