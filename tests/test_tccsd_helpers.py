"""Solver-side helpers: the oracle and the host functions against golden vectors produced by
the reference's own make_no / semi_canonicalize and its literal projection einsums (oracle/gen_golden_tccsd.py); the device
projection of the T2 amplitudes (GPU) against the same goldens and against numpy on larger synthetic amplitudes."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import qio_oracle as O


@pytest.fixture(scope='module')
def gold():
    return dict(np.load(f'{GOLDEN}/ref_tccsd_helpers.npz'))


class _MF:
    def __init__(self, fock):
        self._f = fock

    def get_fock(self):
        return self._f


def test_oracle_helpers_match_reference_goldens(gold):
    no, r = O.make_no(gold['rdm1'], gold['mo'])
    assert np.allclose(np.abs(no), np.abs(gold['no_full']), atol=1e-12) and np.array_equal(r, gold['r_full'])
    no, r = O.make_no(gold['rdm1'], gold['mo'], gold['sub'].tolist())
    assert np.allclose(np.abs(no), np.abs(gold['no_sub']), atol=1e-12) and np.array_equal(r, gold['r_sub'])
    assert np.array_equal(r, gold['rdm1'])            # the reference's sub-block assignment is a no-op (tccsd.py:238)
    sc = O.semi_canonicalize(gold['fock'], gold['mo'], int(gold['nocc']))
    assert np.allclose(np.abs(sc), np.abs(gold['semican']), atol=1e-12)
    assert np.allclose(O.tccsd_project_t2(gold['t2c'], gold['pocc'], gold['pvir']), gold['t2_up'], atol=1e-13)
    assert np.allclose(O.tccsd_project_t2(gold['t2'], gold['pocc'], gold['pvir'], to_cas=True), gold['t2_down'], atol=1e-13)
    t1, t2 = O.tccsd_tailor(gold['t1'], gold['t2'], gold['t1c'], gold['t2c'], gold['pocc'], gold['pvir'])
    assert np.allclose(t1, gold['t1_tailored'], atol=1e-13) and np.allclose(t2, gold['t2_tailored'], atol=1e-12)
    assert 'ijab,Ii,Jj,Aa,Bb->IJAB' in gold['einsum_strings'].tolist()


def test_host_helpers_match_reference_goldens(gold):
    from qio_b200.solver import tccsd as T
    no, r = T.make_no(gold['rdm1'], gold['mo'])
    assert np.array_equal(no, gold['no_full']) and np.array_equal(r, gold['r_full'])
    no, r = T.make_no(gold['rdm1'], gold['mo'], gold['sub'].tolist())
    assert np.array_equal(no, gold['no_sub']) and np.array_equal(r, gold['r_sub'])
    sc = T.semi_canonicalize(_MF(gold['fock']), gold['mo'], int(gold['nocc']))
    assert np.array_equal(sc, gold['semican'])
    assert np.allclose(T.project_t1(gold['t1c'], gold['pocc'], gold['pvir']), gold['t1_up'], atol=1e-13)
    assert np.allclose(T.project_t1(gold['t1'], gold['pocc'], gold['pvir'], to_cas=True), gold['t1_down'], atol=1e-13)


@pytest.mark.gpu
def test_t2_projection_on_device(gold):
    from qio_b200 import device as dev
    from qio_b200.solver import tccsd as T
    up = T.project_t2(gold['t2c'], gold['pocc'], gold['pvir'])
    assert isinstance(up, np.ndarray) and up.shape == gold['t2_up'].shape and np.allclose(up, gold['t2_up'], atol=1e-12)
    down = T.project_t2(dev.to_device(gold['t2']), dev.to_device(gold['pocc']), dev.to_device(gold['pvir']), to_cas=True)
    assert dev.is_device(down) and np.allclose(dev.to_host(down), gold['t2_down'], atol=1e-12)
    kw = dict(t1new=gold['t1'].copy(), t2new=gold['t2'].copy())
    T.make_tailor_callback(gold['t1c'], gold['t2c'], gold['pocc'], gold['pvir'])(kw)
    assert np.allclose(kw['t1new'], gold['t1_tailored'], atol=1e-12) and np.allclose(kw['t2new'], gold['t2_tailored'], atol=1e-11)
    with pytest.raises(ValueError):
        T.project_t2(gold['t2'], gold['pocc'], gold['pvir'])          # CCSD-space amplitudes with the CAS -> CCSD direction


@pytest.mark.gpu
@pytest.mark.parametrize('o,v,Oc,Vc', [(4, 6, 16, 48), (5, 7, 11, 33), (8, 8, 24, 96)])
def test_t2_projection_shapes_of_real_problems(o, v, Oc, Vc):
    """Cr2-like shapes: small CAS (o, v), larger CCSD space (O, V); the TMA path (rows % 16 == 0, even n_out) and the
    cp.async fallback both occur."""
    from qio_b200.solver import tccsd as T
    rng = np.random.default_rng(o * 1000 + Vc)
    pocc, pvir = rng.standard_normal((Oc, o)), rng.standard_normal((Vc, v))
    t2c = rng.standard_normal((o, o, v, v))
    t2 = rng.standard_normal((Oc, Oc, Vc, Vc))
    assert np.allclose(T.project_t2(t2c, pocc, pvir), O.tccsd_project_t2(t2c, pocc, pvir), atol=1e-11, rtol=1e-12)
    assert np.allclose(T.project_t2(t2, pocc, pvir, to_cas=True), O.tccsd_project_t2(t2, pocc, pvir, to_cas=True), atol=1e-10, rtol=1e-12)
