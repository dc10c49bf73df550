"""The C-ABI shared library loads without a GPU, exports every symbol include/*.h declares, and
refuses to compute when there is no device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from triceratops_b200 import _build, _cabi


def _declared():
    text = open(os.path.join(ROOT, "include", "triceratops_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tri_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_build.SO_PATH), "run __graft_entry__.build()"
    assert os.path.dirname(_build.SO_PATH).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_build.SO_PATH)
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.EXPORTS) == names


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_cabi.tri_col) == 16
    assert ctypes.sizeof(_cabi.tri_tp_args) == 8 + 11 * 16 + 8 + 8
    assert ctypes.sizeof(_cabi.tri_eb_args) == 8 + 13 * 16 + 8 + 8
    assert ctypes.sizeof(_cabi.tri_result) == 16 * 8
    # every struct of the binding against the compiler's own sizeof
    sizes = (ctypes.c_int64 * 8)()
    assert _cabi.load().tri_struct_sizes(sizes, 8) == 0
    mine = [ctypes.sizeof(t) for t in (_cabi.tri_col, _cabi.tri_tp_args, _cabi.tri_eb_args,
                                       _cabi.tri_result, _cabi.tri_powerlaw, _cabi.tri_spline,
                                       _cabi.tri_bound_prior, _cabi.tri_sampler_args)]
    assert list(sizes) == mine


def test_calls_before_init_fail_with_state_error():
    lib = _cabi.load()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    assert lib.tri_init(0) == _cabi.TRI_ENODEVICE
    assert b"no CPU fallback" in lib.tri_last_error()
    n = ctypes.c_int32()
    assert lib.tri_sm_count(ctypes.byref(n)) == _cabi.TRI_ESTATE


def test_engine_raises_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from triceratops_b200.engine import Engine
    with pytest.raises(_cabi.TriError):
        Engine(0)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback: without the built CUDA library the package refuses to load."""
    monkeypatch.setenv("TRI_B200_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_cabi, "_lib", None)
    with pytest.raises(ImportError, match="no CPU fallback"):
        _cabi.load()
