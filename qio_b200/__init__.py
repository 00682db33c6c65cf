"""qio_b200 -- B200-native (sm_100a) implementation of the orbital-entropy optimisation core of
schilling-group/QIO behind the reference's own Python API.

    import qio_b200 as qio
    my_qio = qio.QIO(mf, sol, act_space=(n_cas, n_act_e))      # reference qio/__init__.py:3-4
    mo = my_qio.kernel(inactive_indices, mo_coeff, method='nr')

Sub-modules mirror the reference tree: ``qio_b200.qio``, ``qio_b200.entropy``, ``qio_b200.grad.jacobi``,
``qio_b200.grad.gradient``.  ``qio_b200.install_as_qio()`` registers them under the name ``qio`` so
that unmodified user scripts (``import qio``; ``from qio.entropy import get_cost_fqi``) run on the GPU path.

Importing the package loads libqio_b200.so and fails loudly if it has not been built; calling any
routine without a CUDA device raises -- there is no CPU fallback.
"""
from __future__ import annotations

import sys

from . import _lib  # noqa: F401  (loads the CUDA library; ImportError if missing)
from . import device, entropy, grad, qio, solver, synth, tables  # noqa: F401
from ._lib import QioCudaError  # noqa: F401

__version__ = '0.1.0'


def QIO(mf, sol, act_space=None, logger=None, log_file='qio.log'):
    return qio.QIO(mf, sol, act_space, logger, log_file)


def install_as_qio():
    """Make ``import qio`` resolve to this package (drop-in for the reference distribution)."""
    this = sys.modules[__name__]
    sys.modules['qio'] = this
    sys.modules['qio.qio'] = qio
    sys.modules['qio.entropy'] = entropy
    sys.modules['qio.grad'] = grad
    sys.modules['qio.grad.jacobi'] = grad.jacobi
    sys.modules['qio.grad.gradient'] = grad.gradient
    # qio.solver is NOT replaced: the pyscf / block2 wrappers stay the reference's own (qio_b200.solver.tccsd only offers
    # device versions of their dense contractions)
    return this
from . import mutual_info  # noqa: E402,F401  (extension: two-orbital RDM / mutual information)
