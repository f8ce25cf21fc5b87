"""The C-ABI library loads, exports every symbol include/vqacore.h declares, and the ctypes mirror of each
struct has the size the compiler gave it.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from vqa_playground_pytorch_b200 import _lib
    return _lib


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vqacore.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vqa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    L = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libvqacore_sm100a.so does not export %s" % n
    assert set(names) == set(lib.SYMBOLS), set(names) ^ set(lib.SYMBOLS)


def test_struct_layouts(lib):
    L = lib.lib()
    assert L.vqa_abi_version() == lib.ABI_VERSION == 2
    for name, st in lib.STRUCTS.items():
        assert L.vqa_sizeof(name.encode()) == ctypes.sizeof(st), name
    assert L.vqa_sizeof(b"no_such_struct") == 0


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = lib.lib()
    assert L.vqa_device_check() == lib.VQA_ENODEVICE
    assert b"no CPU fallback" in L.vqa_last_error()
    from vqa_playground_pytorch_b200.config import CoR2
    m = CoR2.Model(None, 16)
    with pytest.raises(ValueError, match="CUDA tensor"):
        m({"v": torch.zeros(2, 36, 2048), "q_idxes": torch.zeros(2, 2400)})


def test_bad_arguments_are_rejected(lib):
    L = lib.lib()
    p = lib.LinearFwd()
    p.groups = 99
    assert L.vqa_linear_fwd(ctypes.byref(p), None) == lib.VQA_EINVAL
    assert b"groups" in L.vqa_last_error()
    assert L.vqa_cor2_workspace_bytes(256, 36, 2000) > 0


def test_state_dict_layout_matches_reference():
    from oracle import reasoning_core as rc
    from vqa_playground_pytorch_b200.config import CoR2, ODA
    for cf, name, C in ((CoR2, "CoR2", 2000), (ODA, "ODA", 3000)):
        m = cf.Model(None, C)
        got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
        assert got == rc.param_shapes(name, C)
    m = ODA.Model(None, 3000, num_regions=100)
    assert tuple(m.att.conv_att.conv.weight.shape) == (4, 31000, 1)
