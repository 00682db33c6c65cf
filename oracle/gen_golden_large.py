"""Golden vectors at the HEADLINE sizes, produced by RUNNING THE REAL REFERENCE  --  test infrastructure only.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_large.py [/root/reference] [n100|n86|all]

* ``ref_n100_prefix.npz``: the bench workload (synthetic n = 100 RDMs, ``np.random.seed(0)`` start rotation and
  sweep order, inactive = all) and the first ``PREFIX`` pairs of the first Jacobi cycle.  The reference's own
  ``minimize_orb_corr_jacobi`` would need hours for the whole cycle (1.9 s per accepted rotation), so the loop body
  of qio/grad/jacobi.py:215-231 is driven here pair by pair with the reference's ``jacobi_direct`` /
  ``jacobi_transform`` / ``get_cost_fqi`` doing all the arithmetic.  Stored: angles (NaN = None), costs, U after the
  prefix, gamma, and probes (256 random entries + the diagonal) of the 2-RDM before and after.
* ``ref_n86_gd.npz``: the Cr2 shape (n = 86, nocc = 24): gradient, Hessian and three Newton-Raphson iterations of
  ``minimize_orb_corr_GD`` for both inactive sets.

Nothing from the reference is copied; the GPU box has no /root/reference and reads only the .npz files.
"""
from __future__ import annotations

import logging
import os
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
WHAT = sys.argv[2] if len(sys.argv) > 2 else 'all'
sys.path.insert(0, REF)

# the synthetic generator is plain numpy: load the file directly so that the CUDA library is not needed here
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location('qio_synth', os.path.join(ROOT, 'qio_b200', 'synth.py'))
synth = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(synth)

import qio.qio as ref_qio  # noqa: E402  (the reference)
import qio.entropy as ref_entropy  # noqa: E402
import qio.grad.jacobi as ref_jacobi  # noqa: E402
import qio.grad.gradient as ref_gradient  # noqa: E402
from scipy.linalg import expm  # noqa: E402

logging.getLogger('qio').disabled = True
OUT = os.path.join(ROOT, 'tests', 'golden')
PREFIX = 300


def start_rotation(n):
    """The reference's random start (jacobi.py:197-199) -- consumes rand(n, n), rand() from the global RNG."""
    X = (np.random.rand(n, n) / 2 - 1) / 1 / (1 + 49 * np.random.rand())
    X = X - X.T
    return expm(X)


def rotate(U, gamma, Gamma):
    U_ = np.kron(U, np.eye(2))
    g = np.einsum('ia,jb,ab->ij', U_, U_, gamma, optimize='optimal')          # jacobi.py:201
    G = np.einsum('ia,jb,kc,ld,abcd->ijkl', U, U, U, U, Gamma, optimize='optimal')   # jacobi.py:202
    return g, G


def probes(n, seed):
    return np.random.default_rng(seed).integers(0, n, size=(256, 4))


def make_n100_prefix(n=100):
    t_begin = time.time()
    dm1, dm2 = synth.solver_rdms(n)
    gamma0, Gamma0 = ref_qio.prep_rdm12(dm1, dm2)
    del dm2
    inact = list(range(n))
    np.random.seed(0)
    U = start_rotation(n)
    g, G = rotate(U, gamma0, Gamma0)
    idx = probes(n, 1000 + n)
    k = np.arange(n)
    d = dict(n=n, prefix=PREFIX, Ustart=U.copy(), gamma_rot=g.copy(), probe_idx=idx,
             Gamma0_probe=Gamma0[tuple(idx.T)], Gamma_rot_probe=G[tuple(idx.T)], Gamma_rot_diag=G[k, k, k, k],
             cost_rot=ref_entropy.get_cost_fqi(g, G, inact))
    del Gamma0
    orb_list = np.arange(0, n)
    np.random.shuffle(orb_list)                     # jacobi.py:213-214
    d['order'] = orb_list.copy()
    pairs, thetas, costs = [], [], []
    done = 0
    for a in range(n):                              # jacobi.py:215-231, truncated after PREFIX pairs
        i = orb_list[a]
        for b in range(a):
            j = orb_list[b]
            if done >= PREFIX:
                break
            t = ref_jacobi.jacobi_direct(i, j, g, G, inact)
            pairs.append((int(i), int(j)))
            if t is not None:
                g, G = ref_jacobi.jacobi_transform(g, G, i, j, t)
                V = np.eye(n)
                V[i, i] = V[j, j] = np.cos(t)
                V[i, j] = np.sin(t)
                V[j, i] = -np.sin(t)
                U = V @ U                           # jacobi_step(t, i, j) @ U, jacobi.py:230
                thetas.append(float(t))
                costs.append(ref_entropy.get_cost_fqi(g, G, inact))
            else:
                thetas.append(np.nan)
                costs.append(np.nan)
            done += 1
            if done % 25 == 0:
                print(f'  n100 prefix: {done}/{PREFIX} pairs, {time.time() - t_begin:.0f} s', flush=True)
        if done >= PREFIX:
            break
    d.update(pairs=np.array(pairs, dtype=np.int32), theta=np.array(thetas), cost_after=np.array(costs),
             U_after=U, gamma_after=g, Gamma_after_probe=G[tuple(idx.T)], Gamma_after_diag=G[k, k, k, k],
             cost_final=ref_entropy.get_cost_fqi(g, G, inact))
    path = os.path.join(OUT, 'ref_n100_prefix.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB', f'{time.time() - t_begin:.0f} s')


def make_n86_gd(n=86, seed=14, iters=3):
    dm1, dm2 = synth.solver_rdms(n)
    gamma0, Gamma0 = ref_qio.prep_rdm12(dm1, dm2)
    del dm2
    np.random.seed(seed)
    U = start_rotation(n)
    g, G = rotate(U, gamma0, Gamma0)
    idx = probes(n, 1000 + n)
    k = np.arange(n)
    sets = {'all': list(range(n)), 'partial': [0, 1] + list(range(10, n))}
    active = np.array([q for q in range(n) if q not in sets['partial']])
    d = dict(n=n, seed=seed, Ustart=U, gamma_rot=g, probe_idx=idx, Gamma_rot_probe=G[tuple(idx.T)],
             Gamma_rot_diag=G[k, k, k, k], active=active, gd_iters=iters)
    for name, inact in sets.items():
        d[f'inactive_{name}'] = np.array(inact)
        d[f'cost_{name}'] = ref_entropy.get_cost_fqi(g, G, inact)
        d[f'grad_{name}'] = ref_gradient.FQI_grad(g, G, inact, active)
        d[f'hess_{name}'] = ref_gradient.FQI_hess(g, G, inact, active)
        Ugd, gg, GG, mu = ref_gradient.minimize_orb_corr_GD(g, G, inact, active, thresh=1e-4, max_cycle=iters,
                                                            step_size=1., level_shift=1e-3)
        d[f'gd_U_{name}'] = Ugd
        d[f'gd_gamma_{name}'] = gg
        d[f'gd_cost_{name}'] = ref_entropy.get_cost_fqi(gg, GG, inact)
        d[f'gd_Gdiag_{name}'] = GG[k, k, k, k]
        d[f'gd_Gamma_probe_{name}'] = GG[tuple(idx.T)]
        print(f'  n86 gd {name}: cost {d[f"cost_{name}"]:.6f} -> {d[f"gd_cost_{name}"]:.6f}', flush=True)
    path = os.path.join(OUT, 'ref_n86_gd.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB')


if __name__ == '__main__':
    if WHAT in ('n86', 'all'):
        make_n86_gd()
    if WHAT in ('n100', 'all'):
        make_n100_prefix()
