import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
REFERENCE = os.environ.get('QIO_REFERENCE', '/root/reference')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(n):
    return dict(np.load(os.path.join(GOLDEN, f'ref_n{n}.npz'), allow_pickle=False))


@pytest.fixture(scope='session', params=[6, 12, 28])
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope='session')
def golden6():
    return load_golden(6)


@pytest.fixture(scope='session')
def golden12():
    return load_golden(12)


@pytest.fixture(scope='session')
def golden28():
    return load_golden(28)


def rebuild_inputs(g):
    """(gamma_rot, Gamma_rot) of a golden case; n = 28 stores no full tensor, so it is rebuilt with the
    oracle's rotation from the deterministic synthetic RDMs and checked against the stored probes."""
    from oracle import qio_oracle as O
    from qio_b200 import synth
    n = int(g['n'])
    if 'Gamma_rot' in g:
        return g['gamma_rot'], g['Gamma_rot']
    dm1, dm2 = synth.solver_rdms(n)
    gamma0, Gamma0 = O.prep_rdm12(dm1, dm2)
    G = O.rotate_rdm2(g['Ustart'], Gamma0)
    idx = tuple(g['probe_idx'].T)
    assert np.allclose(G[idx], g['Gamma_rot_probe'], atol=1e-13, rtol=0)
    return g['gamma_rot'], G


def prepared_inputs(g):
    """(gamma0, Gamma0) before the random rotation."""
    from oracle import qio_oracle as O
    from qio_b200 import synth
    if 'Gamma0' in g:
        return g['gamma0'], g['Gamma0']
    dm1, dm2 = synth.solver_rdms(int(g['n']))
    return O.prep_rdm12(dm1, dm2)
