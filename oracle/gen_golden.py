"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE  --  test infrastructure only.

Run in the build container, where the reference tree is mounted read-only:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [/root/reference]

The reference (schilling-group/QIO, pure Python) is imported from the given path; nothing from
it is copied.  Inputs are the deterministic synthetic RDMs of ``qio_b200.synth`` (the small ones
are also stored so the fixtures do not depend on the generator).  The GPU box has no
/root/reference, so the parity tests read only the committed .npz files.
"""
from __future__ import annotations

import logging
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from qio_b200 import synth  # noqa: E402

import qio.qio as ref_qio  # noqa: E402  (the reference)
import qio.entropy as ref_entropy  # noqa: E402
import qio.grad.jacobi as ref_jacobi  # noqa: E402
import qio.grad.gradient as ref_gradient  # noqa: E402

logging.getLogger('qio').disabled = True       # the reference has malformed log calls
OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def inactive_sets(n):
    full = list(range(n))
    if n >= 12:
        partial = [0, 1] + list(range(10, n))      # the QICAS split of examples/c2_dmrg_qicas.py:47-48
    else:
        partial = [0] + list(range(3, n))
    return {'all': full, 'partial': partial}


def generic_basis(n, seed):
    """Prepared RDMs rotated by the reference's own random start (so that pairs are coupled)."""
    dm1, dm2 = synth.solver_rdms(n)
    gamma, Gamma = ref_qio.prep_rdm12(dm1, dm2)
    np.random.seed(seed)
    X = (np.random.rand(n, n) / 2 - 1) / (1 + 49 * np.random.rand())
    X = X - X.T
    from scipy.linalg import expm
    U = expm(X)
    U_ = np.kron(U, np.eye(2))
    g = np.einsum('ia,jb,ab->ij', U_, U_, gamma, optimize='optimal')
    G = np.einsum('ia,jb,kc,ld,abcd->ijkl', U, U, U, U, Gamma, optimize='optimal')
    return dm1, dm2, gamma, Gamma, U, g, G


def theta_or_nan(t):
    return np.nan if t is None else float(t)


def make_case(n, seed, store_full, pair_sample, jacobi_cycles, gd_iters):
    dm1, dm2, gamma0, Gamma0, Ustart, g, G = generic_basis(n, seed)
    d = dict(n=n, seed=seed, Ustart=Ustart, gamma_rot=g)
    if store_full:
        d.update(dm1=dm1, dm2=dm2, gamma0=gamma0, Gamma0=Gamma0, Gamma_rot=G)
    else:
        idx = np.random.default_rng(seed).integers(0, n, size=(256, 4))
        d.update(probe_idx=idx, Gamma0_probe=Gamma0[tuple(idx.T)], Gamma_rot_probe=G[tuple(idx.T)],
                 Gamma0_sum=Gamma0.sum(), Gamma_rot_diag=G[np.arange(n), np.arange(n), np.arange(n), np.arange(n)])
    sets = inactive_sets(n)
    active = np.array([k for k in range(n) if k not in sets['partial']])
    d['active'] = active
    for name, inact in sets.items():
        d[f'inactive_{name}'] = np.array(inact)
        d[f'cost_{name}'] = ref_entropy.get_cost_fqi(g, G, inact)
        d[f'grad_{name}'] = ref_gradient.FQI_grad(g, G, inact, active)
        d[f'hess_{name}'] = ref_gradient.FQI_hess(g, G, inact, active)
        pairs = [(i, j) for i in range(n) for j in range(n) if i != j]
        if pair_sample is not None:
            rng = np.random.default_rng(seed + 1)
            pairs = [pairs[k] for k in rng.choice(len(pairs), size=pair_sample, replace=False)]
        pairs = [(i, j) for (i, j) in pairs if (i in inact or j in inact)]
        thetas = [theta_or_nan(ref_jacobi.jacobi_direct(i, j, g, G, inact)) for (i, j) in pairs]
        d[f'ls_pairs_{name}'] = np.array(pairs)
        d[f'ls_theta_{name}'] = np.array(thetas)
        probe_t = np.array([0.0, 0.01, 0.3, 1.0, 1.5707, 2.2, 3.13])
        d[f'cost_probe_theta'] = probe_t
        d[f'cost_probe_{name}'] = np.array([[ref_jacobi.jacobi_cost(t, i, j, g, G, inact) for t in probe_t]
                                            for (i, j) in pairs[:16]])
    # one Givens transform
    i, j, t = 1, n - 2, 0.4321
    gt, Gt = ref_jacobi.jacobi_transform(g, G, i, j, t)
    d.update(tr_ijt=np.array([i, j, t]), tr_gamma=gt)
    if store_full:
        d['tr_Gamma'] = Gt
    else:
        d['tr_Gamma_probe'] = Gt[tuple(d['probe_idx'].T)]
        d['tr_Gamma_diag'] = Gt[np.arange(n), np.arange(n), np.arange(n), np.arange(n)]
    # reorder_occ
    P, s_val, occ = ref_jacobi.reorder_occ(g.copy(), G.copy())
    d.update(ro_P=P, ro_s=s_val, ro_occ=occ)
    d['rf_P'] = ref_jacobi.reorder_fast(g, G, len(active), 1)
    rots, ncl, V = ref_jacobi.reorder(g, G, len(active))
    d.update(r_rot=np.array(rots).reshape(-1, 3), r_nclosed=ncl, r_V=V)
    # Newton-Raphson ("GD") driver: a few iterations, as QIO.kernel calls it (qio/qio.py:138-140)
    for name, inact in sets.items():
        U, gg, GG, mu = ref_gradient.minimize_orb_corr_GD(g, G, inact, active, thresh=1e-4,
                                                           max_cycle=gd_iters, step_size=1., level_shift=1e-3)
        d[f'gd_U_{name}'] = U
        d[f'gd_gamma_{name}'] = gg
        d[f'gd_cost_{name}'] = ref_entropy.get_cost_fqi(gg, GG, inact)
        d[f'gd_Gdiag_{name}'] = GG[np.arange(n), np.arange(n), np.arange(n), np.arange(n)]
        if store_full:
            d[f'gd_Gamma_{name}'] = GG
    # Jacobi driver from the *un-rotated* prepared RDMs (it draws its own random start)
    for name, inact in sets.items():
        np.random.seed(seed + 100)
        rots, U, gg, GG = ref_jacobi.minimize_orb_corr_jacobi(gamma0, Gamma0, inact, jacobi_cycles)
        d[f'jac_seed'] = seed + 100
        d[f'jac_cycles'] = jacobi_cycles
        d[f'jac_rot_{name}'] = np.array(rots).reshape(-1, 3)
        d[f'jac_U_{name}'] = U
        d[f'jac_gamma_{name}'] = gg
        d[f'jac_cost_{name}'] = ref_entropy.get_cost_fqi(gg, GG, inact)
        d[f'jac_Gdiag_{name}'] = GG[np.arange(n), np.arange(n), np.arange(n), np.arange(n)]
        if store_full:
            d[f'jac_Gamma_{name}'] = GG
    path = os.path.join(OUT, f'ref_n{n}.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


def make_shannon_vectors():
    rng = np.random.default_rng(3)
    specs = [np.array([1.0, 0.0, 0.0, 0.0]), np.array([0.25] * 4), np.array([0.7, 0.2, 0.1, 0.0]),
             np.array([0.5, 0.5 - 1e-9, -1e-9 + 1e-9, 1e-9]), np.array([0.9, 0.1, -1e-9, 1e-9])]
    specs += [rng.dirichlet(np.ones(4)) for _ in range(8)]
    vals = []
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for s in specs:
            vals.append(ref_entropy.shannon(s))
    np.savez_compressed(os.path.join(OUT, 'ref_shannon.npz'), specs=np.array(specs), values=np.array(vals))
    print('wrote ref_shannon.npz')


if __name__ == '__main__':
    make_shannon_vectors()
    make_case(6, seed=11, store_full=True, pair_sample=None, jacobi_cycles=2, gd_iters=4)
    make_case(12, seed=12, store_full=True, pair_sample=None, jacobi_cycles=1, gd_iters=4)
    make_case(28, seed=13, store_full=False, pair_sample=60, jacobi_cycles=1, gd_iters=3)
