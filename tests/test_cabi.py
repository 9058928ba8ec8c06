"""CPU: the C-ABI shared library builds, loads and exports every symbol include/*.h declares
(no compute calls without a GPU), and refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'iodine_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(iodine_[a-z0-9_]+)\s*\(', src)))


def test_library_builds_and_exports_every_declared_symbol():
    from iodine_b200 import _cabi, build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n
    assert sorted(_cabi.EXPORTS) == names
    assert _cabi.load().iodine_abi_version() == 1


def test_struct_layout_matches_header():
    from iodine_b200 import _cabi
    assert ctypes.sizeof(_cabi.IodineShape) == 18 * 4
    assert ctypes.sizeof(_cabi.IodineWeights) == (4 * _cabi.MAX_LAYERS + 14) * 8


def test_plan_create_rejects_bad_shapes_without_gpu_work():
    from iodine_b200 import _cabi
    from iodine_b200.engine import shape_from_arch
    from oracle import arch as A
    lib = _cabi.load()
    s = shape_from_arch(A.arch_by_name('tiny'), 2)
    s.K = 17
    plan = ctypes.c_void_p()
    assert lib.iodine_plan_create(ctypes.byref(s), ctypes.byref(plan)) != 0
    assert b'K' in lib.iodine_last_error()
    s.K, s.dec_chan = 3, 48
    assert lib.iodine_plan_create(ctypes.byref(s), ctypes.byref(plan)) != 0
    # null handles are refused with a message, not dereferenced
    assert lib.iodine_plan_set_comm(None, None, 0, 1) != 0 and b'plan' in lib.iodine_last_error()
    n = ctypes.c_size_t()
    assert lib.iodine_plan_workspace_bytes(None, ctypes.byref(n)) != 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only check')
def test_no_cpu_fallback():
    from iodine_b200 import _cabi
    from iodine_b200.modeling.iodine import IODINE
    from oracle import arch as A
    m = IODINE(A.arch_by_name('tiny'))
    with pytest.raises(_cabi.IodineError):
        m.reconstruct(torch.rand(1, 3, 16, 16))
    with pytest.raises(_cabi.IodineError):
        m(torch.rand(1, 3, 16, 16))
