"""The CPU oracle against golden vectors produced by the real reference (oracle/gen_golden.py)."""
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, prepared_inputs, rebuild_inputs
from oracle import qio_oracle as O
from qio_b200 import synth


def test_shannon_vectors():
    d = np.load(f'{GOLDEN}/ref_shannon.npz')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for s, v in zip(d['specs'], d['values']):
            assert O.shannon(s) == v


def test_shannon_side_effects():
    with pytest.warns(UserWarning):
        O.shannon([0.5, 0.5, -1e-3, 1e-3])
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        O.shannon([0.5, 0.5, -1e-9, 1e-9])          # tiny negative: silent
        O.shannon([1.5, -1e-9])                      # negative present: the >1 check is skipped
    with pytest.raises(ValueError):
        O.shannon([1.5, 0.1])


def test_prep_rdm12_bit_exact(golden6, golden12):
    for g in (golden6, golden12):
        gamma, Gamma = O.prep_rdm12(g['dm1'], g['dm2'])
        assert np.array_equal(gamma, g['gamma0'])
        assert np.array_equal(Gamma, g['Gamma0'])


def test_prep_rdm12_n28_probes(golden28):
    g = golden28
    dm1, dm2 = synth.solver_rdms(28)
    _, Gamma = O.prep_rdm12(dm1, dm2)
    assert np.array_equal(Gamma[tuple(g['probe_idx'].T)], g['Gamma0_probe'])


def test_cost_and_pair_cost(golden):
    g = golden
    gam, Gam = rebuild_inputs(g)
    tol = 0 if 'Gamma_rot' in g else 1e-12
    for name in ('all', 'partial'):
        inact = g[f'inactive_{name}'].tolist()
        assert abs(O.get_cost_fqi(gam, Gam, inact) - g[f'cost_{name}']) <= tol
        pairs = g[f'ls_pairs_{name}'][:16]
        for (i, j), row in zip(pairs, g[f'cost_probe_{name}']):
            for t, v in zip(g['cost_probe_theta'], row):
                assert abs(O.jacobi_cost(t, int(i), int(j), gam, Gam, inact) - v) <= tol
            grid = O.jacobi_cost_grid(g['cost_probe_theta'], int(i), int(j), gam, Gam, inact)
            assert np.allclose(grid, row, atol=tol, rtol=0)


def test_line_search_thetas(golden):
    g = golden
    gam, Gam = rebuild_inputs(g)
    for name in ('all', 'partial'):
        inact = g[f'inactive_{name}'].tolist()
        for (i, j), t_ref in zip(g[f'ls_pairs_{name}'], g[f'ls_theta_{name}']):
            t = O.jacobi_direct(int(i), int(j), gam, Gam, inact)
            if np.isnan(t_ref):
                assert t is None
            else:
                assert t == t_ref


def test_jacobi_transform(golden):
    g = golden
    gam, Gam = rebuild_inputs(g)
    i, j, t = int(g['tr_ijt'][0]), int(g['tr_ijt'][1]), float(g['tr_ijt'][2])
    g2, G2 = O.jacobi_transform(gam, Gam, i, j, t)
    assert np.allclose(g2, g['tr_gamma'], atol=1e-14, rtol=0)
    if 'tr_Gamma' in g:
        assert np.allclose(G2, g['tr_Gamma'], atol=1e-14, rtol=0)
    else:
        assert np.allclose(G2[tuple(g['probe_idx'].T)], g['tr_Gamma_probe'], atol=1e-13, rtol=0)


def test_grad_hess(golden):
    g = golden
    gam, Gam = rebuild_inputs(g)
    for name in ('all', 'partial'):
        inact = g[f'inactive_{name}'].tolist()
        assert np.allclose(O.FQI_grad(gam, Gam, inact), g[f'grad_{name}'], atol=1e-11, rtol=1e-11)
        assert np.allclose(O.FQI_hess(gam, Gam, inact), g[f'hess_{name}'], atol=1e-9, rtol=1e-11)


def test_reorder(golden6, golden12):
    for g in (golden6, golden12):
        gam, Gam = g['gamma_rot'], g['Gamma_rot']
        P, s, occ = O.reorder_occ(gam, Gam)
        assert np.array_equal(P, g['ro_P']) and np.array_equal(s, g['ro_s']) and np.array_equal(occ, g['ro_occ'])
        assert np.array_equal(O.reorder_fast(gam, Gam, len(g['active']), 1), g['rf_P'])
        rots, ncl, V = O.reorder(gam, Gam, len(g['active']))
        assert np.array_equal(np.array(rots).reshape(-1, 3), g['r_rot']) and ncl == int(g['r_nclosed'])
        assert np.array_equal(V, g['r_V'])


def test_newton_driver(golden6, golden12):
    for g in (golden6, golden12):
        for name in ('all', 'partial'):
            inact = g[f'inactive_{name}'].tolist()
            U, gg, GG, mu = O.minimize_orb_corr_GD(g['gamma_rot'], g['Gamma_rot'], inact, g['active'], thresh=1e-4,
                                                   max_cycle=4, step_size=1., level_shift=1e-3)
            assert mu == 2.
            assert np.allclose(U, g[f'gd_U_{name}'], atol=1e-9, rtol=0)
            assert np.allclose(GG, g[f'gd_Gamma_{name}'], atol=1e-9, rtol=0)
            assert abs(O.get_cost_fqi(gg, GG, inact) - g[f'gd_cost_{name}']) < 1e-9


def test_jacobi_driver(golden6, golden12):
    for g in (golden6, golden12):
        gamma0, Gamma0 = prepared_inputs(g)
        for name in ('all', 'partial'):
            inact = g[f'inactive_{name}'].tolist()
            np.random.seed(int(g['jac_seed']))
            rots, U, gg, GG = O.minimize_orb_corr_jacobi(gamma0, Gamma0, inact, int(g['jac_cycles']))
            ref = g[f'jac_rot_{name}']
            got = np.array(rots).reshape(-1, 3)
            assert got.shape == ref.shape
            assert np.array_equal(got, ref)                    # pair order and angles bit-exact
            assert np.allclose(U, g[f'jac_U_{name}'], atol=1e-12, rtol=0)
            assert np.allclose(GG, g[f'jac_Gamma_{name}'], atol=1e-12, rtol=0)
            assert abs(O.get_cost_fqi(gg, GG, inact) - g[f'jac_cost_{name}']) < 1e-12
