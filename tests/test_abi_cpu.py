"""CPU-side checks: the C-ABI library loads and exports every symbol include/scsfm.h declares,
and the host-side argument logic of the wrappers (no kernel launches)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "scsfm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scsfm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for must in ("scsfm_pairwise_fwd", "scsfm_pairwise_bwd", "scsfm_smooth_fwd", "scsfm_smooth_bwd",
                 "scsfm_inverse_warp2_fwd", "scsfm_inverse_warp2_bwd", "scsfm_last_error", "scsfm_version"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from scsfm import lib
    if not os.path.exists(lib.LIB_PATH):
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "sc-sfmlearner-release_b200", "build.sh")])
    dll = ctypes.CDLL(lib.LIB_PATH)
    missing = [s for s in _declared_symbols() if not hasattr(dll, s)]
    assert not missing, missing
    assert lib.load().scsfm_version() >= 100
    assert lib.load().scsfm_pairwise_stats_bytes(4, 4) == 4 * (8 + 48) * 8


def test_argument_errors_are_reported_without_a_gpu():
    from scsfm import lib
    L = lib.load()
    rc = L.scsfm_pairwise_fwd(None, 0, None, 1, 16, 16, 0, 0, None, None, None, None)
    assert rc == -1
    assert b"njobs" in L.scsfm_last_error()
    with pytest.raises(ValueError):
        lib.check(rc, "scsfm_pairwise_fwd")


def test_host_side_helpers():
    import torch
    from scsfm import lib, loss_ops
    assert loss_ops._shift_of(256, 256, "d") == 0
    assert loss_ops._shift_of(256, 32, "d") == 3
    with pytest.raises(ValueError):
        loss_ops._shift_of(256, 100, "d")
    assert loss_ops._flags(1, 1, 1) == 7
    assert loss_ops._flags(2, 1, 0) == 2          # the reference's `== True` test: only 1 enables a term
    assert loss_ops._flags(True, False, True) == 5
    with pytest.raises(ValueError):
        loss_ops._padding("reflection")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lib.dev_f32(torch.zeros(1, 3, 8, 8), "x")
