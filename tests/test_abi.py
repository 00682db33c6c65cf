"""The C-ABI library loads on a CPU-only box and exports every symbol include/qio_b200.h declares;
no compute call is made here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'qio_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(qio_b200_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert 'qio_b200_prep_rdm12' in syms and 'qio_b200_jacobi_cycle' in syms and len(syms) >= 15


def test_library_exports_every_declared_symbol():
    from qio_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    bound = set(_lib.SIGNATURES) | {'qio_b200_last_error'}
    assert set(declared_symbols()) == bound, 'ctypes SIGNATURES out of sync with the header'


def test_struct_layouts_match():
    from qio_b200 import _lib
    assert _lib.LS_DTYPE.itemsize == 80
    assert _lib.LS_DTYPE.fields['theta'][1] == 16 and _lib.LS_DTYPE.fields['cos_theta'][1] == 56
    assert ctypes.sizeof(_lib.ThetaTables) == 8 + 5 * 8
    assert ctypes.sizeof(_lib.SweepOptions) == 4 * 4 + 8 + 8 + 8 and _lib.SweepOptions.blocks_out.offset == 16
    assert ctypes.sizeof(_lib.Peers) == 4 * 4 + 8
    assert _lib.lib.qio_b200_abi_version() == 2


def test_library_keeps_no_process_state():
    """The header promises a stateless ABI: no environment switches and no launch counters in the sources."""
    csrc = os.path.join(ROOT, 'qio_b200', 'csrc')
    for f in sorted(os.listdir(csrc)):
        if f.endswith(('.cu', '.cuh')):
            src = open(os.path.join(csrc, f)).read()
            assert 'getenv' not in src, f'{f} reads the environment'
            assert 'launch_counter' not in src, f'{f} keeps a launch counter'


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import qio_b200
    from qio_b200.entropy import get_cost_fqi
    with pytest.raises(qio_b200.QioCudaError):
        get_cost_fqi(np.zeros((4, 4)), np.zeros((2, 2, 2, 2)), [0, 1])
    with pytest.raises(qio_b200.QioCudaError):
        qio_b200.qio.prep_rdm12(np.eye(2), np.zeros((2, 2, 2, 2)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'qio_b200')
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(base, f)).read()
                assert 'oracle' not in src.replace('no oracle', ''), f'{f} mentions the oracle'
