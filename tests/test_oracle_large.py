"""The oracle against golden vectors the REAL reference produced at the headline sizes (oracle/gen_golden_large.py):
the n = 100 bench workload (start rotation + the first pairs of the first Jacobi cycle) and the Cr2 shape n = 86
(gradient, Hessian, Newton-Raphson driver).  CPU only; sized to stay within about a minute."""
import numpy as np

from conftest import GOLDEN
from oracle import qio_oracle as O
from qio_b200 import synth

PREFIX_ON_CPU = 100     # the golden holds 300 pairs; the GPU test checks all of them


def test_oracle_reproduces_reference_sweep_prefix_n100():
    gold = np.load(f'{GOLDEN}/ref_n100_prefix.npz')
    n = int(gold['n'])
    gamma0, Gamma0 = synth.prepared_rdms(n)
    idx = tuple(gold['probe_idx'].T)
    assert np.allclose(Gamma0[idx], gold['Gamma0_probe'], atol=1e-15, rtol=0)      # built directly, not through prep_rdm12
    # the RNG protocol of the driver: rand(n, n), rand(), shuffle under seed 0 (jacobi.py:197,214)
    np.random.seed(0)
    U = O.random_start(n)
    order = np.arange(n)
    np.random.shuffle(order)
    assert np.array_equal(U, gold['Ustart']) and np.array_equal(order, gold['order'])
    pairs = O.sweep_pairs(order, list(range(n)))
    assert np.array_equal(np.array(pairs[:len(gold['pairs'])]), gold['pairs'])
    g, G = O.rotate_rdm1(U, gamma0), O.rotate_rdm2(U, Gamma0)
    assert np.allclose(G[idx], gold['Gamma_rot_probe'], atol=1e-14, rtol=0)
    inact = list(range(n))
    assert abs(O.get_cost_fqi(g, G, inact) - gold['cost_rot']) < 1e-12
    for (i, j), t_ref, c_ref in list(zip(gold['pairs'], gold['theta'], gold['cost_after']))[:PREFIX_ON_CPU]:
        t = O.jacobi_direct(int(i), int(j), g, G, inact)
        assert (t is None and np.isnan(t_ref)) or t == t_ref           # angles bit-exact, None where the reference has None
        if t is not None:
            g, G = O.jacobi_transform_inplace(g, G, int(i), int(j), t)
    k = int(np.nonzero(~np.isnan(gold['theta'][:PREFIX_ON_CPU]))[0][-1])
    assert abs(O.get_cost_fqi(g, G, inact) - gold['cost_after'][k]) < 1e-11


def test_jacobi_transform_inplace_is_jacobi_transform():
    rng = np.random.default_rng(3)
    n = 7
    gamma, Gamma = rng.standard_normal((2 * n, 2 * n)), rng.standard_normal((n, n, n, n))
    g1, G1 = O.jacobi_transform(gamma, Gamma, 5, 2, 0.77)
    G2 = Gamma.copy()
    g2, G2b = O.jacobi_transform_inplace(gamma, G2, 5, 2, 0.77)
    assert G2b is G2 and np.array_equal(G1, G2) and np.array_equal(g1, g2)


def test_oracle_reproduces_reference_newton_n86():
    gold = np.load(f'{GOLDEN}/ref_n86_gd.npz')
    n = int(gold['n'])
    gamma0, Gamma0 = synth.prepared_rdms(n)
    g, G = O.rotate_rdm1(gold['Ustart'], gamma0), O.rotate_rdm2(gold['Ustart'], Gamma0)
    idx = tuple(gold['probe_idx'].T)
    assert np.allclose(G[idx], gold['Gamma_rot_probe'], atol=1e-14, rtol=0)
    inact = gold['inactive_partial'].tolist()
    assert abs(O.get_cost_fqi(g, G, inact) - gold['cost_partial']) < 1e-11
    assert np.allclose(O.FQI_grad(g, G, inact), gold['grad_partial'], atol=1e-10, rtol=1e-10)
    assert np.allclose(O.FQI_hess(g, G, inact), gold['hess_partial'], atol=1e-9, rtol=1e-10)
    U, gg, GG, mu = O.minimize_orb_corr_GD(g, G, inact, gold['active'], thresh=1e-4, max_cycle=int(gold['gd_iters']),
                                           step_size=1., level_shift=1e-3)
    # the Newton step divides by ~1e-4 and exponentiates: rounding differences of the rotation are amplified
    assert np.allclose(U, gold['gd_U_partial'], atol=1e-7, rtol=0)
    assert abs(O.get_cost_fqi(gg, GG, inact) - gold['gd_cost_partial']) < 1e-7
