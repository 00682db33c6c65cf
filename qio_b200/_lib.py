"""ctypes binding of the C ABI in include/qio_b200.h (libqio_b200.so, built in-tree by
``__graft_entry__.build()`` / ``make -C qio_b200/csrc``).

There is NO CPU fallback: if the shared library is missing the import fails, and if no CUDA
device is present every compute call raises ``QioCudaError``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('QIO_B200_LIB') or os.path.join(_HERE, 'csrc', 'libqio_b200.so')   # override: A/B builds

BLOCK_DOUBLES = 24
FLAG_WARN_NEGATIVE = 1
FLAG_RAISE_GT1 = 2

# numpy view of struct qio_b200_ls_result (80 bytes)
LS_DTYPE = np.dtype([
    ('coarse_index', np.int32), ('fine_index', np.int32), ('flags', np.int32), ('accepted', np.int32),
    ('theta', np.float64), ('cost', np.float64), ('cost0', np.float64), ('coarse_margin', np.float64),
    ('fine_margin', np.float64), ('cos_theta', np.float64), ('sin_theta', np.float64), ('reserved', np.float64),
])
assert LS_DTYPE.itemsize == 80


class QioCudaError(RuntimeError):
    """Raised when the CUDA library reports an error (or no device is available)."""


class ThetaTables(ctypes.Structure):
    """struct qio_b200_theta_tables"""
    _fields_ = [('n_coarse', ctypes.c_int32), ('fine_stride', ctypes.c_int32),
                ('coarse_theta', ctypes.c_void_p), ('coarse_cs', ctypes.c_void_p),
                ('fine_len', ctypes.c_void_p), ('fine_theta', ctypes.c_void_p),
                ('fine_cs', ctypes.c_void_p)]


class SweepOptions(ctypes.Structure):
    """struct qio_b200_sweep_options (all zero = defaults)"""
    _fields_ = [('engine', ctypes.c_int32), ('timers', ctypes.c_int32), ('timeout_ms', ctypes.c_uint32),
                ('startup_timeout_ms', ctypes.c_uint32), ('blocks_out', ctypes.c_void_p), ('trace_out', ctypes.c_void_p),
                ('trace_first_pair', ctypes.c_int32), ('reserved', ctypes.c_int32)]


class Peers(ctypes.Structure):
    """struct qio_b200_peers"""
    _fields_ = [('rank', ctypes.c_int32), ('nranks', ctypes.c_int32), ('epoch', ctypes.c_uint32),
                ('reserved', ctypes.c_uint32), ('mailboxes', ctypes.c_void_p)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"qio_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; "
        f"g.build()'` or `make -C qio_b200/csrc`. There is no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

_p = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
_ll = ctypes.c_longlong

# name -> argtypes; every entry of include/qio_b200.h (tests/test_abi.py cross-checks the header)
SIGNATURES = {
    'qio_b200_abi_version': ([], _i),
    'qio_b200_device_info': ([_p, _p, _p], _i),
    'qio_b200_prep_rdm12': ([_p, _p, _p, _p, _i, _i, _p], _i),
    'qio_b200_selftest_div6': ([ctypes.c_ulonglong, ctypes.c_ulonglong, _p, _p], _i),
    'qio_b200_orbital_diagonals': ([_p, _p, _i, _i, _i, _p, _i, _p, _p], _i),
    'qio_b200_orbital_entropies': ([_p, _i, _p, _p, _p, _p, _p], _i),
    'qio_b200_shannon': ([_p, _ll, _p, _p], _i),
    'qio_b200_pair_blocks': ([_p, _p, _i, _i, _i, _p, _i, _p, _p], _i),
    'qio_b200_fqi_grad_hess': ([_p, _i, _p, _p, _p, _p], _i),
    'qio_b200_pair_grad_hess': ([_p, _p, _i, _p, _p, _p, _p], _i),
    'qio_b200_fqi_grad_hess_fused': ([_p, _p, _i, _p, _p, _p, _p], _i),
    'qio_b200_pair_cost': ([_p, _p, _p, _p, _i, _p, _p, _p, _p], _i),
    'qio_b200_pair_linesearch': ([_p, _p, _i, _p, ctypes.POINTER(ThetaTables), _p, _i, _p], _i),
    'qio_b200_givens_apply': ([_p, _p, _i, _i, _i, _d, _d, _i, _p], _i),
    'qio_b200_rotate_rdm1': ([_p, _p, _p, _i, _p], _i),
    'qio_b200_rotate_rdm2': ([_p, _p, _p, _i, _p], _i),
    'qio_b200_rotate_axis': ([_p, _p, _p, _i, _ll, _ll, _p], _i),
    'qio_b200_rotate_axis_batched': ([_p, _p, _p, _i, _ll, _ll, _i, _ll, _ll, _p], _i),
    'qio_b200_contract_axis': ([_p, _p, _p, _i, _ll, _i, _p], _i),
    'qio_b200_contract_axis_cyclic': ([_p, _ll, _p, _p, _i, _i, _ll, _p], _i),
    'qio_b200_contract_axis_partial': ([_p, _ll, _p, _p, _i, _i, _ll, _p], _i),
    'qio_b200_sweep_workspace_bytes': ([_i], ctypes.c_size_t),
    'qio_b200_sweep_engine': ([_i, _i], _i),
    'qio_b200_sweep_reset': ([_p, _i, _p], _i),
    'qio_b200_mailbox_bytes': ([_i], ctypes.c_size_t),
    'qio_b200_sweep_cycle': ([_p, _p, _p, _p, _i, _i, _i, _p, _i, _p, ctypes.POINTER(ThetaTables), _p, _p, _p,
                              ctypes.POINTER(Peers), ctypes.POINTER(SweepOptions), _p], _i),
    'qio_b200_deferred_diagonals': ([_p, _p, _p, _i, _i, _i, _p, _i, _p, _p], _i),
    'qio_b200_ipc_alloc': ([ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p], _i),
    'qio_b200_ipc_open': ([ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)], _i),
    'qio_b200_ipc_close': ([_p], _i),
    'qio_b200_ipc_free': ([_p], _i),
    'qio_b200_ipc_clear': ([_p, ctypes.c_size_t, _p], _i),
    'qio_b200_ll_stress_bytes': ([_i], ctypes.c_size_t),
    'qio_b200_ll_stress': ([ctypes.POINTER(Peers), _i, _i, _p, _p], _i),
    'qio_b200_zero_offspin_if_rotated': ([_p, _i, _p, _p], _i),
    'qio_b200_pair_entropy': ([_p, _p, _i, _p, _p, _p, _p], _i),
    'qio_b200_mutual_information': ([_p, _p, _p, _i, _i, _p, _p], _i),
    'qio_b200_jacobi_workspace_bytes': ([_i, _i], ctypes.c_size_t),
    'qio_b200_jacobi_cycle': ([_p, _p, _p, _i, _p, _i, _p, ctypes.POINTER(ThetaTables), _p, _p, _p, _p], _i),
}

lib.qio_b200_last_error.argtypes = []
lib.qio_b200_last_error.restype = ctypes.c_char_p
for _name, (_args, _res) in SIGNATURES.items():
    _fn = getattr(lib, _name)     # AttributeError here == the library is stale: fail loudly
    _fn.argtypes = _args
    _fn.restype = _res


def last_error() -> str:
    return lib.qio_b200_last_error().decode('utf-8', 'replace')


def check(rc: int, what: str = '') -> None:
    if rc != 0:
        raise QioCudaError(f"{what or 'qio_b200'} failed (rc={rc}): {last_error()}")
