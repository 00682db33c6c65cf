"""Gradient layer -- the public functions of the reference's ``qio/grad/gradient.py`` (same names,
positional signatures, return values), executed by the CUDA library.

Per-pair primitives (gamma_grad_diag ... orb_entr_hess, gradient.py:13-120) and the all-pairs
matrices FQI_grad / FQI_hess (:123-146) come from one device pass over the pair blocks;
minimize_orb_corr_GD (:155-232) keeps the RDMs resident on the device and leaves only the n x n
work on the host (level shift, element-wise Newton step, ``scipy.linalg.expm``).
"""
from __future__ import annotations

import logging

import numpy as np
from scipy.linalg import expm

from .. import device as dev
from ..entropy import get_cost_fqi, raise_or_warn

logger = logging.getLogger('qio')


# ---- small per-pair helpers (host indexing of a handful of entries; kept for API parity) -----------
def _entry(x, *idx):
    v = x[idx]
    return float(v.item()) if dev.is_device(x) else float(v)


def gamma_grad_diag(gamma, i, j):
    """d gamma_kk / d theta at 0 for the rotation (i, j); the down entries copy the up value
    (gradient.py:13-23)."""
    dg = np.zeros(len(gamma))
    up = 2 * _entry(gamma, 2 * i, 2 * j)
    dg[2 * i], dg[2 * j], dg[2 * i + 1], dg[2 * j + 1] = up, -up, up, -up
    return dg


def gamma_hess_diag(gamma, i, j):
    """d2 gamma_kk / d theta2 at 0 (gradient.py:25-32)."""
    ddg = np.zeros(len(gamma))
    diff = _entry(gamma, 2 * i, 2 * i) - _entry(gamma, 2 * j, 2 * j)
    ddg[2 * i], ddg[2 * j] = -2 * diff, 2 * diff
    ddg[2 * i + 1], ddg[2 * j + 1] = ddg[2 * i], ddg[2 * j]
    return ddg


def Gamma_grad_diag(Gamma, i, j):
    """d Gamma_kkkk / d theta at 0 (gradient.py:34-42; the [i,j,j,j] entry is counted twice there)."""
    G = lambda *idx: _entry(Gamma, *idx)
    dG = np.zeros(len(Gamma))
    dG[i] = G(j, i, i, i) + G(i, j, i, i) + G(i, i, j, i) + G(i, i, i, j)
    dG[j] = -G(i, j, j, j) - G(i, j, j, j) - G(j, j, i, j) - G(j, j, j, i)
    return dG


def Gamma_hess_diag(Gamma, i, j):
    """d2 Gamma_kkkk / d theta2 at 0 (gradient.py:44-50)."""
    G = lambda *idx: _entry(Gamma, *idx)
    ddG = np.zeros(len(Gamma))
    mixed = 2 * (G(j, j, i, i) + G(j, i, j, i) + G(j, i, i, j) + G(i, j, j, i) + G(i, j, i, j) + G(i, i, j, j))
    ddG[i] = -4 * G(i, i, i, i) + mixed
    ddG[j] = -4 * G(j, j, j, j) + mixed
    return ddG


def act_occ_grad(gamma, i, j, active_indices):
    dg = gamma_grad_diag(gamma, i, j)
    return np.array([dg[2 * k] + dg[2 * k + 1] for k in set(active_indices).intersection({i, j})])


def act_occ_hess(gamma, i, j, active_indices):
    ddg = gamma_hess_diag(gamma, i, j)
    return np.array([ddg[2 * k] + ddg[2 * k + 1] for k in set(active_indices).intersection({i, j})])


def _pair_grad_hess_members(gamma, Gamma, i, j, inactive_indices):
    """Per-member (dS, ddS) of one pair from the device: one evaluation per inactive member of {i, j}."""
    n = int(Gamma.shape[0])
    g, G = dev.to_device(gamma), dev.to_device(Gamma)
    pair = dev.to_device(np.array([[int(i), int(j)]], dtype=np.int32), dev.I32)
    blocks = dev.pair_blocks(g, G, n, pair)
    out = []
    for k in set(inactive_indices).intersection({i, j}):
        g1, g2 = dev.pair_grad_hess(blocks, pair, dev.inactive_mask([k], n))
        out.append((float(g1.item()), float(g2.item())))
    return out


def orb_entr_grad(gamma, Gamma, i, j, inactive_indices, active_indices):
    """dS_k/dtheta at 0 for each inactive k in {i, j} (gradient.py:62-81)."""
    return np.array([m[0] for m in _pair_grad_hess_members(gamma, Gamma, i, j, inactive_indices)])


def orb_entr_hess(gamma, Gamma, i, j, inactive_indices, active_indices):
    """d2S_k/dtheta2 at 0 for each inactive k in {i, j} (gradient.py:93-120)."""
    return np.array([m[1] for m in _pair_grad_hess_members(gamma, Gamma, i, j, inactive_indices)])


# ---- all-pairs matrices --------------------------------------------------------------------------
def grad_hess_device(g, G, n, mask):
    """(dX, ddX) CUDA tensors for device-resident RDMs."""
    return dev.fqi_grad_hess_fused(g, G, n, mask)


def FQI_grad(gamma, Gamma, inactive_indices, active_indices, mu=0.):
    """Antisymmetric (n, n) matrix of d cost / d theta_ij (gradient.py:123-133)."""
    n = int(Gamma.shape[0])
    dX, _ = grad_hess_device(dev.to_device(gamma), dev.to_device(Gamma), n, dev.inactive_mask(inactive_indices, n))
    return dev.like_input(dX, Gamma)


def FQI_hess(gamma, Gamma, inactive_indices, active_indices, mu=0.):
    """Symmetric (n, n) matrix of d2 cost / d theta_ij2 (gradient.py:135-146)."""
    n = int(Gamma.shape[0])
    _, ddX = grad_hess_device(dev.to_device(gamma), dev.to_device(Gamma), n, dev.inactive_mask(inactive_indices, n))
    return dev.like_input(ddX, Gamma)


def FQI_display(gamma, Gamma, inactive_indices, verbose=True):
    cost_ = get_cost_fqi(gamma, Gamma, inactive_indices)
    logger.info('FQI cost: %s', cost_)
    return cost_


def newton_step(grad, hess, step_size):
    """Host n x n arithmetic of gradient.py:204-208: level shift = smallest |hess|>1e-12 entry + 1e-4
    (the ``level_shift`` argument of the driver is overwritten there), X = -grad/(hess - shift)*step
    where |denominator| > 1e-12, else 0."""
    no = grad.shape[0]
    min_hess = np.min((hess[np.abs(hess) > 1e-12]))
    level_shift = min_hess + 1e-4
    denom = hess - level_shift * np.ones((no, no))
    return -np.divide(grad, denom, out=np.zeros_like(grad), where=np.abs(denom) > 1e-12) * step_size


def minimize_orb_corr_GD(gamma_, Gamma_, inactive_indices, active_indices, step_size=0.1, thresh=1e-4,
                         max_cycle=100, level_shift=1e-3, mu=2., mu_rate=1., target_n_ae=None, noise=1e-3):
    """Diagonal-Hessian Newton iteration on the orbital-entropy cost (gradient.py:155-232).

    Returns (U_tot, gamma, Gamma, mu); every step is accepted, iteration stops when the cost change
    is within ``thresh`` or after ``max_cycle`` iterations.  RDMs stay on the device throughout."""
    no = int(Gamma_.shape[0])
    g = dev.to_device(gamma_).clone()
    G = dev.to_device(Gamma_).clone()
    work = dev.empty(no, no, no, no)
    mask = dev.inactive_mask(inactive_indices, no)
    inds = dev.to_device(np.asarray(inactive_indices, dtype=np.int32), dev.I32)

    def cost_of(gd, Gd):
        spec, _, _, summary = dev.orbital_entropies(dev.orbital_diagonals(gd, Gd, no, inds), want_spec=True)
        total, _, _, flags = dev.to_host(summary)
        if int(flags):
            raise_or_warn(int(flags), dev.to_host(spec))
        return np.float64(total)

    U_tot = expm(np.zeros((no, no)))
    n = 0
    cost_old = cost_of(g, G)
    delta_cost = np.inf
    _ = 2 * active_indices            # gradient.py:198 requires an ndarray here
    while np.abs(delta_cost) > thresh and n < max_cycle:
        n += 1
        dX, ddX = grad_hess_device(g, G, no, mask)
        grad, hess = dev.to_host(dX), dev.to_host(ddX)
        X = newton_step(grad, hess, step_size)
        U = expm(X)
        U_tot = np.matmul(U, U_tot)
        Ud = dev.to_device(U)
        g = dev.rotate_rdm1(Ud, g, no)
        dev.rotate_rdm2_(Ud, G, no, work)
        cost = cost_of(g, G)
        delta_cost = cost_old - cost
        cost_old = cost
        logger.info("iteration:" + str(n) + " max |grad| = " + str(np.max(abs(grad))) + " cost = " + str(cost))

    return U_tot, dev.like_input(g, gamma_), dev.like_input(G, Gamma_), mu
