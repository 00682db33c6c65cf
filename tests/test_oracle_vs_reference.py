"""Differential test: the oracle against the REAL reference imported from /root/reference
(build container only; skipped where the reference tree is absent, e.g. on the GPU box)."""
import logging
import os
import sys

import numpy as np
import pytest

from conftest import REFERENCE
from oracle import qio_oracle as O
from qio_b200 import synth

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, 'qio')), reason='reference tree not mounted')


@pytest.fixture(scope='module')
def ref():
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'qio' or k.startswith('qio.')}
    sys.path.insert(0, REFERENCE)
    try:
        import qio.entropy as e
        import qio.grad.gradient as gr
        import qio.grad.jacobi as j
        import qio.qio as q
        logging.getLogger('qio').disabled = True
        yield dict(q=q, e=e, j=j, g=gr)
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k == 'qio' or k.startswith('qio.')]:
            del sys.modules[k]
        sys.modules.update(saved)
        logging.getLogger('qio').disabled = False


@pytest.fixture(scope='module')
def case():
    n = 10
    dm1, dm2 = synth.solver_rdms(n, seed=77)
    gamma0, Gamma0 = O.prep_rdm12(dm1, dm2)
    np.random.seed(3)
    U = O.random_start(n)
    return dict(n=n, dm1=dm1, dm2=dm2, gamma0=gamma0, Gamma0=Gamma0, g=O.rotate_rdm1(U, gamma0),
                G=O.rotate_rdm2(U, Gamma0), sets=[list(range(n)), [0, 1] + list(range(5, n))], act=np.array([2, 3, 4]))


def test_prep(ref, case):
    a = ref['q'].prep_rdm12(case['dm1'], case['dm2'])
    assert np.array_equal(a[0], case['gamma0']) and np.array_equal(a[1], case['Gamma0'])


def test_costs_and_search(ref, case):
    g, G = case['g'], case['G']
    for inact in case['sets']:
        assert ref['e'].get_cost_fqi(g, G, inact) == O.get_cost_fqi(g, G, inact)
        for (i, j) in [(3, 1), (0, 7), (9, 2), (4, 5)]:
            for t in (0, 0.01, 0.77, 2.9):
                assert ref['j'].jacobi_cost(t, i, j, g, G, inact) == O.jacobi_cost(t, i, j, g, G, inact)
            tr = ref['j'].jacobi_direct(i, j, g, G, inact)
            assert tr == O.jacobi_direct(i, j, g, G, inact) == O.jacobi_direct_scalar(i, j, g, G, inact)


def test_transform(ref, case):
    a = ref['j'].jacobi_transform(case['g'], case['G'], 2, 8, 1.234)
    for impl in (O.jacobi_transform, O.jacobi_transform_dense):
        b = impl(case['g'], case['G'], 2, 8, 1.234)
        assert np.allclose(a[0], b[0], atol=1e-15, rtol=0) and np.allclose(a[1], b[1], atol=1e-15, rtol=0)


def test_grad_hess_and_newton(ref, case):
    g, G, act = case['g'], case['G'], case['act']
    for inact in case['sets']:
        assert np.array_equal(ref['g'].FQI_grad(g, G, inact, act), O.FQI_grad(g, G, inact, act))
        assert np.array_equal(ref['g'].FQI_hess(g, G, inact, act), O.FQI_hess(g, G, inact, act))
    a = ref['g'].minimize_orb_corr_GD(g, G, case['sets'][1], act, step_size=1., max_cycle=3)
    b = O.minimize_orb_corr_GD(g, G, case['sets'][1], act, step_size=1., max_cycle=3)
    for x, y in zip(a[:3], b[:3]):      # the Newton step divides by ~1e-4: rounding differences are amplified
        assert np.allclose(x, y, atol=1e-8, rtol=0)


def test_jacobi_driver(ref, case):
    np.random.seed(9)
    a = ref['j'].minimize_orb_corr_jacobi(case['gamma0'], case['Gamma0'], case['sets'][1], 2)
    np.random.seed(9)
    b = O.minimize_orb_corr_jacobi(case['gamma0'], case['Gamma0'], case['sets'][1], 2)
    assert np.array_equal(np.array(a[0]), np.array(b[0]))
    for x, y in zip(a[1:], b[1:]):
        assert np.allclose(x, y, atol=1e-13, rtol=0)


def test_reorder(ref, case):
    g, G = case['g'], case['G']
    for x, y in zip(ref['j'].reorder_occ(g, G), O.reorder_occ(g, G)):
        assert np.array_equal(x, y)
    assert np.array_equal(ref['j'].reorder_fast(g, G, 3, 2), O.reorder_fast(g, G, 3, 2))
    a, b = ref['j'].reorder(g, G, 3), O.reorder(g, G, 3)
    assert a[0] == b[0] and a[1] == b[1] and np.array_equal(a[2], b[2])
