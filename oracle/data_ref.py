"""Load the UNMODIFIED reference datasets (``/root/reference/lib/data/clevr.py``, ``dsprite.py``) for the input
pipeline tests and fixtures.  TEST INFRASTRUCTURE -- only usable where ``/root/reference`` is mounted.

The two files import ``skimage.io`` (and ``h5py``), which this image does not have, and use ``np.float``, which
numpy removed.  They are loaded from where they lie with three shims that do not touch their logic:
``skimage.io.imread`` -> a PIL decode (same uint8 array for PNG files), an empty ``h5py`` module (imported, never
used), and ``np.float = float`` for the duration of a call.
"""
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np

from .ref_loader import REFERENCE_ROOT


def data_reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'lib', 'data', 'clevr.py'))


def _pil_imread(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im)


def _shims():
    if 'skimage' not in sys.modules:
        sk, io = types.ModuleType('skimage'), types.ModuleType('skimage.io')
        io.imread = _pil_imread
        sk.io = io
        sys.modules['skimage'], sys.modules['skimage.io'] = sk, io
    if 'h5py' not in sys.modules:
        sys.modules['h5py'] = types.ModuleType('h5py')


def load_reference_dataset_module(name):
    """name: 'clevr' or 'dsprite' -> the module object of ``lib/data/<name>.py`` (not through the package
    ``__init__``, which would pull MNIST and the rest in)."""
    _shims()
    path = os.path.join(REFERENCE_ROOT, 'lib', 'data', name + '.py')
    spec = importlib.util.spec_from_file_location('iodine_reference_data_' + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def legacy_numpy():
    had = hasattr(np, 'float')
    if not had:
        np.float = float
    try:
        yield
    finally:
        if not had:
            del np.float
