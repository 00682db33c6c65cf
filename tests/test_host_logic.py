"""Host-side logic of the product (no GPU): pair sequences, angle tables, Newton step, reordering."""
import numpy as np

from oracle import qio_oracle as O
from qio_b200 import synth, tables
from qio_b200.grad import gradient as G_
from qio_b200.grad import jacobi as J_


def test_sweep_pairs_match_reference_loop_order():
    rng = np.random.default_rng(0)
    for n, inact in ((7, list(range(7))), (12, [0, 1] + list(range(10, 12))), (5, [3]), (4, [])):
        order = rng.permutation(n)
        got = J_.sweep_pairs(order, inact)
        want = np.array(O.sweep_pairs(order, inact), dtype=np.int32).reshape(-1, 2)
        assert got.dtype == np.int32 and np.array_equal(got, want)


def test_angle_tables_are_the_reference_grids():
    t = tables.host_tables()
    coarse = O.coarse_grid()
    assert np.array_equal(t['coarse_theta'], coarse) and len(coarse) == 314
    assert np.array_equal(t['coarse_cs'][:, 0], np.cos(coarse)) and np.array_equal(t['coarse_cs'][:, 1], np.sin(coarse))
    for r in (0, 1, 57, 156, 313):
        f = O.fine_grid(coarse[r])
        L = int(t['fine_len'][r])
        assert L == len(f) and np.array_equal(t['fine_theta'][r, :L], f)
        assert np.array_equal(t['fine_cs'][r, :L, 0], np.cos(f)) and np.array_equal(t['fine_cs'][r, :L, 1], np.sin(f))
    # the t_opt = 0 -> 0.01 fallback is coarse point 0
    t_opt = 0
    t_opt += 0.01
    assert np.array_equal(O.fine_grid(t_opt), O.fine_grid(coarse[0]))
    assert set(t['fine_len'].tolist()) <= {200, 201} and t['fine_stride'] == 201


def test_newton_step_matches_oracle():
    rng = np.random.default_rng(1)
    g = rng.standard_normal((9, 9)); g = g - g.T
    h = rng.standard_normal((9, 9)); h = h + h.T; np.fill_diagonal(h, 0)
    assert np.array_equal(G_.newton_step(g, h, 0.7), O.newton_step(g, h, 0.7))


def test_reorder_on_numpy_inputs_matches_oracle():
    gamma, Gamma = synth.prepared_rdms(9, seed=5)
    np.random.seed(2)
    U = O.random_start(9)
    g, G = O.rotate_rdm1(U, gamma), O.rotate_rdm2(U, Gamma)
    for a, b in zip(J_.reorder_occ(g, G), O.reorder_occ(g, G)):
        assert np.array_equal(a, b)
    assert np.array_equal(J_.reorder_fast(g, G, 3, 2), O.reorder_fast(g, G, 3, 2))


def test_small_derivative_helpers_match_oracle():
    gamma, Gamma = synth.prepared_rdms(6, seed=8)
    np.random.seed(4)
    U = O.random_start(6)
    g, G = O.rotate_rdm1(U, gamma), O.rotate_rdm2(U, Gamma)
    dg, ddg, dG, ddG = O.pair_derivative_inputs(g, G, 4, 1)
    assert G_.gamma_grad_diag(g, 4, 1)[8] == dg[4][0] and G_.gamma_grad_diag(g, 4, 1)[3] == dg[1][1]
    assert G_.gamma_hess_diag(g, 4, 1)[8] == ddg[4][0]
    assert G_.Gamma_grad_diag(G, 4, 1)[4] == dG[4] and G_.Gamma_grad_diag(G, 4, 1)[1] == dG[1]
    assert G_.Gamma_hess_diag(G, 4, 1)[1] == ddG[1]


def test_synthetic_rdms_are_consistent():
    n = 8
    dm1, dm2 = synth.solver_rdms(n)
    nel = 2 * synth.default_nocc(n)
    assert abs(np.trace(dm1) - nel) < 1e-12 and abs(np.einsum('iijj->', dm2) - nel * (nel - 1)) < 1e-10
    g0, G0 = O.prep_rdm12(dm1, dm2)
    g1, G1 = synth.prepared_rdms(n)
    assert np.allclose(g0, g1, atol=1e-15) and np.allclose(G0, G1, atol=1e-15)
    assert np.allclose(G0, G0.transpose(1, 0, 3, 2)) and np.allclose(G0, G0.transpose(3, 2, 1, 0))
