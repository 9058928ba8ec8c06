"""ctypes binding of the C ABI in ``include/iodine_b200.h`` (no torch types cross it).

The product path has NO fallback: if the shared library is missing or a call fails, an
exception is raised.  ``iodine_b200.build.build()`` compiles the library in-tree.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libiodine_b200.so')

MAX_LAYERS = 8
ABI_VERSION = 2
FP32, BF16, TF32, FP16 = 0, 1, 2, 3
PRECISIONS = {'fp32': FP32, 'bf16': BF16, 'fp16': FP16, 'tf32': TF32}

EXPORTS = [
    'iodine_abi_version', 'iodine_last_error', 'iodine_plan_create', 'iodine_plan_destroy',
    'iodine_plan_workspace_bytes', 'iodine_plan_set_workspace', 'iodine_plan_set_weights',
    'iodine_init_state', 'iodine_refine_step', 'iodine_elbo', 'iodine_encode', 'iodine_decode',
    'iodine_reconstruct', 'iodine_reconstruct_host', 'iodine_reconstruct_host_async', 'iodine_debug_read',
    'iodine_plan_launch_count', 'iodine_plan_profile', 'iodine_plan_profile_read', 'iodine_ari',
    'iodine_plan_set_comm', 'iodine_plan_last_elbo_image0', 'iodine_evaluate_host', 'iodine_evaluate_host_async',
    'iodine_plan_train_workspace_bytes', 'iodine_plan_set_train_workspace', 'iodine_train_step',
]


class IodineShape(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        'B', 'K', 'L', 'H', 'W', 'T', 'img_c', 'dec_layers', 'dec_chan', 'dec_k',
        'ref_layers', 'ref_chan', 'ref_k', 'ref_stride', 'mlp_units', 'layernorm')] + [
        ('sigma', C.c_float), ('precision', C.c_int32), ('slot_ranks', C.c_int32), ('slot_rank', C.c_int32)]


_FP = C.c_void_p


class IodineWeights(C.Structure):
    _fields_ = [
        ('dec_w', _FP * MAX_LAYERS), ('dec_b', _FP * MAX_LAYERS),
        ('dec_out_w', _FP), ('dec_out_b', _FP),
        ('ref_w', _FP * MAX_LAYERS), ('ref_b', _FP * MAX_LAYERS),
        ('mlp_w', _FP), ('mlp_b', _FP),
        ('lstm_w_ih', _FP), ('lstm_w_hh', _FP), ('lstm_b_ih', _FP), ('lstm_b_hh', _FP),
        ('mean_w', _FP), ('mean_b', _FP), ('logvar_w', _FP), ('logvar_b', _FP),
        ('init_mean', _FP), ('init_logvar', _FP),
    ]


class IodineGrads(C.Structure):
    """gradient outputs of iodine_train_step: the same fields as IodineWeights"""
    _fields_ = list(IodineWeights._fields_)


class IodineError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the shared library; raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IodineError(
            'libiodine_b200.so is not built (%s). Run `python -m iodine_b200.build`; '
            'there is no CPU/PyTorch fallback for the refinement loop.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
    lib.iodine_abi_version.restype = i32
    lib.iodine_last_error.restype = C.c_char_p
    sig = {
        'iodine_plan_create': [C.POINTER(IodineShape), C.POINTER(vp)],
        'iodine_plan_destroy': [vp],
        'iodine_plan_workspace_bytes': [vp, C.POINTER(sz)],
        'iodine_plan_set_workspace': [vp, vp, sz],
        'iodine_plan_set_weights': [vp, C.POINTER(IodineWeights), vp],
        'iodine_init_state': [vp, vp, vp, vp, vp, vp],
        'iodine_refine_step': [vp] * 10,
        'iodine_elbo': [vp] * 7,
        'iodine_encode': [vp] * 7,
        'iodine_decode': [vp] * 6,
        'iodine_reconstruct': [vp] * 9,
        'iodine_reconstruct_host': [vp] * 9,
        'iodine_reconstruct_host_async': [vp] * 9,
        'iodine_debug_read': [vp, C.c_char_p, vp, sz, C.POINTER(sz), vp],
        'iodine_plan_launch_count': [vp, C.POINTER(C.c_uint64)],
        'iodine_plan_profile': [vp, i32],
        'iodine_plan_profile_read': [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)],
        'iodine_ari': [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp],
        'iodine_plan_set_comm': [vp, vp, i32, i32],
        'iodine_plan_last_elbo_image0': [vp] * 5,
        'iodine_evaluate_host': [vp] * 8,
        'iodine_evaluate_host_async': [vp] * 8,
        'iodine_plan_train_workspace_bytes': [vp, C.POINTER(sz)],
        'iodine_plan_set_train_workspace': [vp, vp, sz],
        'iodine_train_step': [vp, vp, vp, i32, C.POINTER(IodineGrads), vp, vp, vp],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    if lib.iodine_abi_version() != ABI_VERSION:
        raise IodineError('ABI version mismatch: %d' % lib.iodine_abi_version())
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise IodineError(load().iodine_last_error().decode('utf-8', 'replace'))
