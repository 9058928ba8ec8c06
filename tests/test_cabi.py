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
    assert _cabi.load().iodine_abi_version() == 2


def test_struct_layout_matches_header():
    from iodine_b200 import _cabi
    assert ctypes.sizeof(_cabi.IodineShape) == 20 * 4
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


def test_plain_c_host_compiles_and_links(tmp_path):
    """examples/host_c.c: the ABI is usable from C99 with nothing but the header and the .so (run on a B200:
    INTEGRATION.md); here it is compiled with -Wall -Werror and linked, not executed."""
    import shutil
    import subprocess
    if not shutil.which('gcc'):
        pytest.skip('no gcc')
    from iodine_b200 import _cabi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_cabi.LIB_PATH)
    cuda = next((d for d in ('/usr/local/cuda/lib64', '/usr/local/cuda/targets/x86_64-linux/lib')
                 if os.path.exists(os.path.join(d, 'libcudart.so'))), None)
    if cuda is None:
        pytest.skip('no libcudart to link against')
    _cabi.load()
    out = str(tmp_path / 'host_c')
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', os.path.join(root, 'examples', 'host_c.c'),
                        '-I', os.path.join(root, 'include'), '-L', libdir, '-liodine_b200', '-L', cuda, '-lcudart',
                        '-Wl,-rpath,' + libdir, '-o', out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(out)
