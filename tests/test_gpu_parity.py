"""GPU parity: the CUDA path (through the Python mirror -> ctypes -> C ABI) against the CPU oracle and
against golden vectors produced by the real reference.  Run on the B200 box: pytest -m gpu.

Tolerances: bit-exact for prep_rdm12, pair order, accepted angles (they are table entries); 1e-10
absolute for entropies / costs (north_star); 1e-10 relative for gradient / Hessian entries, which the
reference divides by spectrum entries that can be tiny."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN, prepared_inputs, rebuild_inputs
from oracle import qio_oracle as O

pytestmark = pytest.mark.gpu

import qio_b200  # noqa: E402
from qio_b200 import device as dev  # noqa: E402
from qio_b200 import synth, tables  # noqa: E402
from qio_b200.entropy import get_cost_fqi, shannon  # noqa: E402
from qio_b200.grad import gradient as GR  # noqa: E402
from qio_b200.grad import jacobi as JA  # noqa: E402
from qio_b200.qio import prep_rdm12  # noqa: E402

ENT_TOL = 1e-10


def rotated_case(n, seed=5, np_seed=1):
    gamma0, Gamma0 = synth.prepared_rdms(n, seed=seed)
    np.random.seed(np_seed)
    U = O.random_start(n)
    return O.rotate_rdm1(U, gamma0), O.rotate_rdm2(U, Gamma0)


# ---- a-1 ------------------------------------------------------------------------------------------
# odd n and n < 16: the cp-style kernel (53+: several 52-wide tiles, ragged last tile); even n >= 16: the TMA kernel (64-wide tile
# pairs combined in place: one tile, exactly one tile, ragged second tile, three tiles)
@pytest.mark.parametrize('n', [1, 2, 6, 12, 16, 18, 31, 32, 33, 40, 52, 53, 54, 60, 64, 66, 80, 130])
def test_prep_rdm12_bit_exact(n):
    rng = np.random.default_rng(n)
    dm1 = rng.standard_normal((n, n))
    dm2 = rng.standard_normal((n, n, n, n))          # no symmetry at all: every element position is checked
    g_ref, G_ref = O.prep_rdm12(dm1, dm2)
    g, G = prep_rdm12(dm1, dm2)
    assert isinstance(G, np.ndarray) and np.array_equal(g, g_ref) and np.array_equal(G, G_ref)
    gd, Gd = prep_rdm12(dev.to_device(dm1), dev.to_device(dm2))
    assert dev.is_device(Gd) and np.array_equal(dev.to_host(Gd), G_ref) and np.array_equal(dev.to_host(gd), g_ref)


def test_div6_three_operation_division_is_ieee_division():
    """The prep kernel divides by 6 with q0 = x * RN(1/6), r = fma(-6, q0, x), q = fma(r, RN(1/6), q0): bit-identical to
    the IEEE division on 2 x 2e9 pseudo-random doubles (all exponents, and the range RDM entries live in)."""
    bad = torch.zeros(1, dtype=torch.int64, device='cuda')
    from qio_b200._lib import check, lib
    check(lib.qio_b200_selftest_div6(2_000_000_000, 12345, bad.data_ptr(), None), 'selftest_div6')
    torch.cuda.synchronize()
    assert int(bad.item()) == 0


def test_prep_rdm12_special_values():
    """Zeros (both signs), subnormals, huge values, infinities and NaNs go through the same combination as numpy's."""
    n = 16
    rng = np.random.default_rng(0)
    dm2 = rng.standard_normal((n, n, n, n))
    flat = dm2.reshape(-1)
    flat[::7] = 0.0
    flat[1::11] = -0.0
    flat[2::13] = 5e-324
    flat[3::17] = 1e-310
    flat[4::19] = 1e305
    flat[5::23] = np.inf
    flat[6::29] = np.nan
    flat[8::31] = 2.0 ** -1000
    with np.errstate(all='ignore'):
        _, G_ref = O.prep_rdm12(np.eye(n), dm2)
    _, G = prep_rdm12(np.eye(n), dm2)
    assert np.array_equal(G, G_ref, equal_nan=True)
    assert np.array_equal(np.signbit(G[~np.isnan(G)]), np.signbit(G_ref[~np.isnan(G_ref)]))


def test_prep_rdm12_golden(golden6, golden12):
    for gold in (golden6, golden12):
        g, G = prep_rdm12(gold['dm1'], gold['dm2'])
        assert np.array_equal(g, gold['gamma0']) and np.array_equal(G, gold['Gamma0'])


def test_prep_rdm12_shard_local():
    n = 10
    rng = np.random.default_rng(0)
    dm2 = rng.standard_normal((n, n, n, n))
    _, G_ref = O.prep_rdm12(np.eye(n), dm2)
    for a0, na in ((0, 3), (3, 4), (7, 3), (5, 0)):
        _, Gs = dev.prep_rdm12(None, dev.to_device(dm2[a0:a0 + na]), n, na)
        assert np.array_equal(dev.to_host(Gs).reshape(na, n, n, n), G_ref[a0:a0 + na])


# ---- a-2 / a-3 --------------------------------------------------------------------------------------
def test_shannon_golden_and_side_effects():
    d = np.load(f'{GOLDEN}/ref_shannon.npz')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for s, v in zip(d['specs'], d['values']):
            assert abs(shannon(s) - v) < 1e-14
    with pytest.warns(UserWarning):
        shannon([0.5, 0.5, -1e-3, 1e-3])
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        shannon([0.5, 0.5, -1e-9, 1e-9])
        shannon([1.5, -1e-9])
    with pytest.raises(ValueError):
        shannon([1.5, 0.1])
    big = np.random.default_rng(1).dirichlet(np.ones(1000))
    assert abs(shannon(big) - O.shannon(big)) < 1e-12
    assert shannon(np.array([1.0, 0.0, 0.0, 0.0])) == 0.0
    assert abs(shannon([0.25] * 4) - np.log(4)) < 1e-15


def test_get_cost_fqi(golden):
    gam, Gam = rebuild_inputs(golden)
    for name in ('all', 'partial'):
        inact = golden[f'inactive_{name}'].tolist()
        assert abs(get_cost_fqi(gam, Gam, inact) - golden[f'cost_{name}']) < ENT_TOL
    Gd, gd = dev.to_device(Gam), dev.to_device(gam)
    assert abs(get_cost_fqi(gd, Gd, golden['inactive_all'].tolist()) - golden['cost_all']) < ENT_TOL


def test_cost_uniform_and_empty():
    n = 6
    gamma = np.eye(2 * n) * 0.5
    Gamma = np.zeros((n, n, n, n)); Gamma[np.arange(n), np.arange(n), np.arange(n), np.arange(n)] = 0.25
    assert abs(get_cost_fqi(gamma, Gamma, list(range(n))) - n * np.log(4)) < 1e-13
    assert get_cost_fqi(gamma, Gamma, []) == 0.0


# ---- a-4 / a-5 --------------------------------------------------------------------------------------
def test_jacobi_cost_golden(golden):
    gam, Gam = rebuild_inputs(golden)
    gd, Gd = dev.to_device(gam), dev.to_device(Gam)
    for name in ('all', 'partial'):
        inact = golden[f'inactive_{name}'].tolist()
        for (i, j), row in zip(golden[f'ls_pairs_{name}'][:8], golden[f'cost_probe_{name}']):
            for t, v in zip(golden['cost_probe_theta'], row):
                assert abs(JA.jacobi_cost(t, int(i), int(j), gd, Gd, inact) - v) < ENT_TOL


@pytest.mark.parametrize('strict', [False, True])
def test_jacobi_direct_golden_angles_bit_exact(golden, strict):
    gam, Gam = rebuild_inputs(golden)
    gd, Gd = dev.to_device(gam), dev.to_device(Gam)
    for name in ('all', 'partial'):
        inact = golden[f'inactive_{name}'].tolist()
        pairs, want = golden[f'ls_pairs_{name}'], golden[f'ls_theta_{name}']
        rec = JA.linesearch_pairs(gd, Gd, pairs, inact, strict_order=strict)
        got = np.where(rec['accepted'] != 0, rec['theta'], np.nan)
        assert np.array_equal(got, want, equal_nan=True)
        # the scalar API on a few pairs
        for (i, j), t_ref in list(zip(pairs, want))[:5]:
            t = JA.jacobi_direct(int(i), int(j), gd, Gd, inact)
            assert (t is None and np.isnan(t_ref)) or t == t_ref


@pytest.mark.parametrize('strict', [False, True])
def test_linesearch_all_pairs_vs_oracle_n40(strict):
    n = 40
    g, G = rotated_case(n)
    inact = [0, 1] + list(range(10, n))
    pairs = [(i, j) for i in range(n) for j in range(i) if (i in inact or j in inact)]
    rec = JA.linesearch_pairs(g, G, pairs, inact, strict_order=strict)
    close = 0
    for (i, j), r in zip(pairs, rec):
        info = O.jacobi_direct(i, j, g, G, inact, details=True)
        if min(info['fine_margin'], info['coarse_gap']) < 1e-13:
            close += 1                       # decision closer than rounding: either answer is legitimate
            continue
        assert r['coarse_index'] == info['coarse_index'] and r['fine_index'] == info['fine_index']
        assert (info['theta'] is None and not r['accepted']) or r['theta'] == info['theta']
        assert abs(r['cost'] - info['cost']) < ENT_TOL and abs(r['cost0'] - info['cost0']) < ENT_TOL
        assert abs(r['fine_margin'] - info['fine_margin']) < 1e-12
    assert close <= 2


def determinant_rdms(n, nocc):
    """Closed-shell determinant in its own orbital basis: Gamma[a,b,c,d] = P[a,d] P[b,c], P = diag(occ)."""
    P = np.diag((np.arange(n) < nocc).astype(float))
    gamma = np.zeros((2 * n, 2 * n))
    gamma[0::2, 0::2] = P
    gamma[1::2, 1::2] = P
    return gamma, np.ascontiguousarray(np.einsum('ad,bc->abcd', P, P))


def test_linesearch_inactive_and_uphill_edge_cases():
    n = 6
    gamma, Gamma = determinant_rdms(n, 3)
    assert get_cost_fqi(gamma, Gamma, list(range(n))) == 0.0
    # neither orbital inactive: the pair cost is identically 0 and no angle is returned
    assert JA.jacobi_cost(0.3, 0, 4, gamma, Gamma, [2]) == 0.0
    assert JA.jacobi_direct(0, 4, gamma, Gamma, [2]) is None
    # occupied-virtual rotation only raises the entropy: None, like the oracle
    assert JA.jacobi_direct(0, 4, gamma, Gamma, list(range(n))) is None
    assert O.jacobi_direct(0, 4, gamma, Gamma, list(range(n))) is None
    v = JA.jacobi_cost(0.5, 0, 4, gamma, Gamma, list(range(n)))
    assert abs(v - O.jacobi_cost(0.5, 0, 4, gamma, Gamma, list(range(n)))) < 1e-14 and v > 0.5


# ---- a-6 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,i,j', [(2, 0, 1), (3, 2, 0), (6, 1, 4), (9, 8, 0), (12, 3, 4), (17, 5, 16), (28, 27, 26)])
def test_jacobi_transform_vs_oracle(n, i, j):
    rng = np.random.default_rng(n)
    gamma = rng.standard_normal((2 * n, 2 * n))      # generic, incl. off-spin entries that must be zeroed
    Gamma = rng.standard_normal((n, n, n, n))        # no symmetry: every orbit class is exercised
    g_ref, G_ref = O.jacobi_transform(gamma, Gamma, i, j, 0.4321)
    g, G = JA.jacobi_transform(gamma, Gamma, i, j, 0.4321)
    assert np.allclose(G, G_ref, atol=1e-13, rtol=0) and np.allclose(g, g_ref, atol=1e-13, rtol=0)
    untouched = np.ones(n, bool); untouched[[i, j]] = False
    sub = np.ix_(untouched, untouched, untouched, untouched)
    assert np.array_equal(G[sub], Gamma[sub])          # entries without i, j are not rewritten


def test_jacobi_transform_golden(golden6, golden12):
    for gold in (golden6, golden12):
        i, j, t = int(gold['tr_ijt'][0]), int(gold['tr_ijt'][1]), float(gold['tr_ijt'][2])
        g, G = JA.jacobi_transform(gold['gamma_rot'], gold['Gamma_rot'], i, j, t)
        assert np.allclose(G, gold['tr_Gamma'], atol=1e-13, rtol=0) and np.allclose(g, gold['tr_gamma'], atol=1e-13, rtol=0)


def test_transform_then_cost_equals_rotated_pair_cost():
    n = 14
    g, G = rotated_case(n)
    inact = list(range(n))
    i, j, t = 9, 2, 0.8
    g2, G2 = JA.jacobi_transform(g, G, i, j, t)
    others = [k for k in inact if k not in (i, j)]
    lhs = get_cost_fqi(g2, G2, inact)
    rhs = JA.jacobi_cost(t, i, j, g, G, inact) + get_cost_fqi(g, G, others)
    assert abs(lhs - rhs) < 1e-12


# ---- a-8 ------------------------------------------------------------------------------------------
def test_fqi_grad_hess_golden(golden):
    gam, Gam = rebuild_inputs(golden)
    for name in ('all', 'partial'):
        inact = golden[f'inactive_{name}'].tolist()
        grad = GR.FQI_grad(gam, Gam, inact, golden['active'])
        hess = GR.FQI_hess(gam, Gam, inact, golden['active'])
        assert np.allclose(grad, golden[f'grad_{name}'], atol=1e-10, rtol=1e-10)
        assert np.allclose(hess, golden[f'hess_{name}'], atol=1e-10, rtol=1e-10)
        assert np.array_equal(grad, -grad.T) and np.array_equal(hess, hess.T)
        assert not grad.diagonal().any() and not hess.diagonal().any()


def test_grad_hess_block_path_and_members():
    n = 11
    g, G = rotated_case(n)
    inact = [0, 1] + list(range(4, n))
    gd, Gd = dev.to_device(g), dev.to_device(G)
    a, b = np.tril_indices(n, -1)
    pairs = dev.to_device(np.stack([a, b], 1).astype(np.int32), dev.I32)
    blocks = dev.pair_blocks(gd, Gd, n, pairs)
    mask = dev.inactive_mask(inact, n)
    dX, ddX = dev.fqi_grad_hess(blocks, n, mask)
    dXf, ddXf = dev.fqi_grad_hess_fused(gd, Gd, n, mask)
    assert torch.equal(dX, dXf) and torch.equal(ddX, ddXf)
    assert np.allclose(dev.to_host(dX), O.FQI_grad(g, G, inact), atol=1e-11, rtol=1e-11)
    # blocks themselves are exact copies
    Gb, gu, gdn = O.pair_block(g, G, 7, 3)
    k = 7 * 6 // 2 + 3
    rec = dev.to_host(blocks[k])
    assert np.array_equal(rec[:16], Gb.ravel()) and np.array_equal(rec[16:20], gu.ravel()) and np.array_equal(rec[20:], gdn.ravel())
    # per-member API, both index orders
    for (i, j) in ((7, 3), (3, 7), (2, 9)):
        want = [O.pair_grad_hess(g, G, i, j, [k_]) for k_ in set(inact).intersection({i, j})]
        got_g = GR.orb_entr_grad(g, G, i, j, inact, None)
        got_h = GR.orb_entr_hess(g, G, i, j, inact, None)
        assert np.allclose(got_g, [w[0] for w in want], atol=1e-11, rtol=1e-11)
        assert np.allclose(got_h, [w[1] for w in want], atol=1e-10, rtol=1e-11)


def test_gradient_is_derivative_of_pair_cost():
    n = 10
    g, G = rotated_case(n)
    inact = list(range(n))
    grad = GR.FQI_grad(g, G, inact, None)
    h = 1e-5
    for (i, j) in ((5, 2), (9, 0)):
        fd = (JA.jacobi_cost(h, i, j, g, G, inact) - JA.jacobi_cost(-h, i, j, g, G, inact)) / (2 * h)
        assert abs(fd - grad[i, j]) < 1e-7


# ---- rotations ------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [2, 5, 12, 16, 17, 28, 34])
def test_rotate_rdm2_vs_oracle(n):
    rng = np.random.default_rng(n)
    Gamma = rng.standard_normal((n, n, n, n))
    gamma = rng.standard_normal((2 * n, 2 * n))
    U = rng.standard_normal((n, n))                      # need not be orthogonal
    Gd = dev.to_device(Gamma).clone()
    Ud = dev.to_device(U)
    dev.rotate_rdm2_(Ud, Gd, n)
    ref = O.rotate_rdm2(U, Gamma)
    assert np.allclose(dev.to_host(Gd), ref, atol=1e-11 * max(1, n), rtol=1e-12)
    gout = dev.rotate_rdm1(Ud, dev.to_device(gamma), n)
    assert np.allclose(dev.to_host(gout), O.rotate_rdm1(U, gamma), atol=1e-11, rtol=1e-12)


# ---- a-10 -----------------------------------------------------------------------------------------
def test_newton_driver_golden(golden6, golden12):
    for gold in (golden6, golden12):
        for name in ('all', 'partial'):
            inact = gold[f'inactive_{name}'].tolist()
            U, gg, GG, mu = GR.minimize_orb_corr_GD(gold['gamma_rot'], gold['Gamma_rot'], inact, gold['active'],
                                                    thresh=1e-4, max_cycle=4, step_size=1., level_shift=1e-3)
            assert mu == 2. and isinstance(GG, np.ndarray)
            # the Newton step divides by ~1e-4 and exponentiates: rounding differences are amplified
            assert np.allclose(U, gold[f'gd_U_{name}'], atol=1e-7, rtol=0)
            assert abs(get_cost_fqi(gg, GG, inact) - gold[f'gd_cost_{name}']) < 1e-8
            # self-consistency at full precision: returned RDMs are the input rotated by the returned U
            assert np.allclose(GG, O.rotate_rdm2(U, gold['Gamma_rot']), atol=1e-10, rtol=0)


# ---- a-7 ------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n,rows', [(6, 216), (12, 100), (17, 60), (28, 28 ** 3)])
def test_contract_axis_keep_order(n, rows):
    rng = np.random.default_rng(n)
    U = rng.standard_normal((n, n))
    lead = rng.standard_normal((n, rows))
    out = dev.contract_axis(dev.to_device(U), dev.to_device(lead), dev.empty(n, rows), n, rows, last_axis=False)
    assert np.allclose(dev.to_host(out), U @ lead, atol=1e-11, rtol=1e-12)
    last = rng.standard_normal((rows, n))
    out = dev.contract_axis(dev.to_device(U), dev.to_device(last), dev.empty(rows, n), n, rows, last_axis=True)
    assert np.allclose(dev.to_host(out), last @ U.T, atol=1e-11, rtol=1e-12)


@pytest.mark.parametrize('n', [5, 12, 28])
def test_deferred_engine_matches_materialized_engine(n):
    """The persistent deferred-axis cycle and the per-pair materialised cycle are independent device
    implementations: same accept / reject pattern and angles, same RDMs after the flush."""
    g, G = rotated_case(n, seed=3, np_seed=4)
    inact = [0, 1] + list(range(3, n))
    order = np.random.default_rng(n).permutation(n)
    pairs = JA.sweep_pairs(order, inact)
    out = {}
    for engine in ('materialized', 'deferred', 'deferred-barrier'):
        gd, Gd, Ud = dev.to_device(g).clone(), dev.to_device(G).clone(), dev.to_device(np.eye(n)).clone()
        st = JA.SweepState(gd, Gd, Ud, n, inact, engine=engine)
        rec1, acc1, cost1, _ = st.cycle(pairs)
        rec2, acc2, cost2, _ = st.cycle(pairs[::-1].copy())          # second cycle on the deferred state
        if engine != 'materialized':
            assert st.last_audit['pairs'] == len(pairs) and not st.last_audit['unexplained']
            assert st.last_summary[24] == (3.0 if engine == 'deferred' else 0.0)      # which kernel really ran
        gg, GG = st.finish()
        out[engine] = (rec1, rec2, cost1, cost2, dev.to_host(gg), dev.to_host(GG), dev.to_host(Ud))
    # the pipelined and the barrier kernel sum the same partial products in different orders: same decisions
    c = out['deferred-barrier']
    for k in (0, 1):
        assert np.array_equal(c[k]['accepted'], out['deferred'][k]['accepted'])
        assert np.array_equal(c[k]['theta'], out['deferred'][k]['theta'], equal_nan=True)
    assert np.allclose(c[5], out['deferred'][5], atol=1e-12, rtol=0)
    a, b = out['materialized'], out['deferred']
    for k in (0, 1):
        assert np.array_equal(a[k]['accepted'], b[k]['accepted'])
        assert np.array_equal(a[k]['theta'], b[k]['theta'], equal_nan=True)
        assert np.allclose(a[k]['cost'], b[k]['cost'], atol=1e-12, rtol=0)
    assert abs(a[2] - b[2]) < 1e-12 and abs(a[3] - b[3]) < 1e-12
    assert np.allclose(a[4], b[4], atol=1e-12, rtol=0) and np.allclose(a[5], b[5], atol=1e-12, rtol=0)
    assert np.allclose(a[6], b[6], atol=1e-12, rtol=0)
    # and both agree with rotating the input by the accumulated U
    assert np.allclose(b[5], O.rotate_rdm2(b[6], G), atol=1e-11, rtol=0)
    assert abs(get_cost_fqi(b[4], b[5], inact) - b[3]) < 1e-12


@pytest.mark.parametrize('n,K', [(6, 1), (6, 2), (6, 3), (6, 7), (12, 5), (12, 30), (13, 9), (28, 40)])
def test_partial_pair_list_pipelined_matches_barrier_kernel(n, K):
    """A launch that visits only the first K pairs leaves orbitals that never entered a pair ("dormant" in the
    pipelined kernel: their rows are skipped by the update and caught up at the end of the launch).  Both sweep
    kernels must leave the same state and records."""
    g, G = rotated_case(n, seed=5, np_seed=6)
    inact = list(range(n))
    order = np.random.default_rng(100 + n).permutation(n)
    pairs = JA.sweep_pairs(order, inact)[:K]
    out = {}
    for eng in ('3', '2'):
        gd, Gd, Ud = dev.to_device(g).clone(), dev.to_device(G).clone(), dev.to_device(np.eye(n)).clone()
        W = dev.empty(n, n)
        dev.sweep_reset(W, n)
        res, summ = dev.sweep_cycle(gd, Gd, Ud, W, n, dev.to_device(pairs, dev.I32), dev.inactive_mask(inact, n),
                                    tables.device_tables(), engine=int(eng))
        dev.check_sweep_summary(dev.to_host(summ))
        out[eng] = (dev.results_to_numpy(res), dev.to_host(Gd), dev.to_host(gd), dev.to_host(Ud), dev.to_host(W))
    assert dev.sweep_engine(n) == 3
    a, b = out['3'], out['2']
    assert np.array_equal(a[0]['accepted'], b[0]['accepted'])
    assert np.array_equal(a[0]['theta'], b[0]['theta'], equal_nan=True)
    assert K < 3 or a[0]['accepted'].sum() > 0
    for k in (1, 2, 3, 4):
        assert np.allclose(a[k], b[k], atol=1e-12, rtol=0)


@pytest.mark.parametrize('engine', ['deferred', 'deferred-barrier', 'materialized'])
def test_jacobi_driver_golden_rotations_bit_exact(golden, engine, monkeypatch):
    """'deferred' = the pipelined kernel (sweep3.cu), 'deferred-barrier' = the same state with one grid barrier per pair
    (sweep2.cu, forced through the options of the call), 'materialized' = one launch per pair on the natural layout."""
    monkeypatch.setattr(JA, 'ENGINE', engine)
    gamma0, Gamma0 = prepared_inputs(golden)
    n = int(golden['n'])
    for name in ('all', 'partial'):
        inact = golden[f'inactive_{name}'].tolist()
        np.random.seed(int(golden['jac_seed']))
        rots, U, gg, GG = JA.minimize_orb_corr_jacobi(gamma0, Gamma0, inact, int(golden['jac_cycles']))
        ref = golden[f'jac_rot_{name}']
        got = np.array(rots).reshape(-1, 3)
        assert got.shape == ref.shape
        assert np.array_equal(got[:, :2], ref[:, :2])           # sweep order and accept / reject pattern
        assert np.array_equal(got[:, 2], ref[:, 2])             # angles (degrees) bit-exact
        assert np.allclose(U, golden[f'jac_U_{name}'], atol=1e-11, rtol=0)
        assert np.allclose(gg, golden[f'jac_gamma_{name}'], atol=1e-11, rtol=0)
        assert abs(get_cost_fqi(gg, GG, inact) - golden[f'jac_cost_{name}']) < ENT_TOL
        k = np.arange(n)
        assert np.allclose(GG[k, k, k, k], golden[f'jac_Gdiag_{name}'], atol=1e-11, rtol=0)
        if f'jac_Gamma_{name}' in golden:
            assert np.allclose(GG, golden[f'jac_Gamma_{name}'], atol=1e-11, rtol=0)


# ---- driver ---------------------------------------------------------------------------------------
class _Mol:
    def __init__(self, nelectron):
        self.nelectron = nelectron


class _MF:
    def __init__(self, n, nelectron):
        self.mo_coeff = np.eye(n)
        self.mol = _Mol(nelectron)
        self.verbose = 0


class _FixedSolver:
    """Your_Solver contract (reference README.md:32-42) returning fixed RDMs."""
    def __init__(self, dm1, dm2):
        self.dm1, self.dm2, self.calls = dm1, dm2, 0

    def kernel(self, mo_coeff):
        self.calls += 1

    def make_rdm1(self):
        return self.dm1

    def make_rdm2(self):
        return self.dm2


def test_qio_kernel_matches_oracle_pipeline(tmp_path):
    import logging
    n = 12
    dm1, dm2 = synth.solver_rdms(n)
    mf = _MF(n, 2 * synth.default_nocc(n))
    q = qio_b200.QIO(mf, _FixedSolver(dm1, dm2), act_space=(4, 2), logger=logging.getLogger('qio_test_quiet'))
    q.max_cycle = 3
    inact = [k for k in range(n) if k not in q.active_indices]
    mo = q.kernel(inact, None, method='NR')
    g0, G0 = O.prep_rdm12(dm1, dm2)
    U, gg, GG, _ = O.minimize_orb_corr_GD(g0, G0, inact, q.active_indices, thresh=1e-4, max_cycle=3, step_size=1.)
    V, s_val, occ = O.reorder_occ(gg, GG)
    assert np.allclose(mo, np.eye(n) @ (V @ U).T, atol=1e-7)
    assert np.allclose(q.occ_num, occ, atol=1e-8) and isinstance(q.Gamma, np.ndarray)
    assert abs(get_cost_fqi(q.gamma, q.Gamma, inact) - O.get_cost_fqi(gg, GG, inact)) < 1e-8
    with pytest.raises(NotImplementedError):
        q.kernel(inact, None, method='bfgs')
    q.keep_on_device = True
    np.random.seed(0)
    q.max_cycle = 1
    q.kernel(inact, None, method='jacobi')
    assert dev.is_device(q.Gamma)
    np.random.seed(0)
    rots, Uj, gj, Gj = O.minimize_orb_corr_jacobi(g0, G0, inact, 1)
    assert abs(get_cost_fqi(q.gamma, q.Gamma, inact) - O.get_cost_fqi(gj, Gj, inact)) < ENT_TOL


def test_reorder_device_inputs(golden12):
    gold = golden12
    gd, Gd = dev.to_device(gold['gamma_rot']), dev.to_device(gold['Gamma_rot'])
    P, s, occ = JA.reorder_occ(gd, Gd)
    assert np.array_equal(P, gold['ro_P']) and np.array_equal(occ, gold['ro_occ']) and np.array_equal(s, gold['ro_s'])
    assert np.array_equal(JA.reorder_fast(gd, Gd, len(gold['active']), 1), gold['rf_P'])
    rots, ncl, V = JA.reorder(gd, Gd, len(gold['active']))
    assert np.array_equal(np.array(rots).reshape(-1, 3), gold['r_rot']) and ncl == int(gold['r_nclosed'])
    assert np.array_equal(V, gold['r_V'])
