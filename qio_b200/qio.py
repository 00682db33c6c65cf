"""Driver layer -- ``QIO`` with the attributes, defaults and ``kernel`` contract of the reference's
``qio/qio.py`` (class QIO :9-148, prep_rdm12 :152-174, get_overlap :177-189).

The solver object (``kernel(mo_coeff)``, ``make_rdm1()``, ``make_rdm2()``; reference README.md:32-42)
stays on the host and is called exactly as the reference calls it.  Everything between
``prep_rdm12`` and ``reorder_occ`` runs on the GPU; the 4-index ``dm2`` is uploaded once, in slabs of
the leading index through pinned staging buffers, overlapped with the prep kernel.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import device as dev
from .grad.gradient import minimize_orb_corr_GD
from .grad.jacobi import minimize_orb_corr_jacobi, reorder_occ


class QIO:
    def __init__(self, mf, sol, act_space=None, logger=None, log_file='qio.log'):
        """
        Attributes (names and defaults as in the reference, qio/qio.py:28-47,76):
            mf: mean-field object (uses mf.mo_coeff, mf.mol.nelectron, mf.verbose)
            sol: solver providing kernel(mo_coeff), make_rdm1(), make_rdm2()
            act_space (tuple): (n_cas, n_act_e)
            max_cycle, thresh, level_shift, step_size, mu, mu_rate, debug, verbose
        Saved results: mo_coeff, gamma, Gamma, s_val, occ_num.
        Extra: keep_on_device (default False) leaves gamma / Gamma as CUDA tensors after kernel().
        Extension (no reference counterpart): with ``compute_mutual_info = True`` kernel() also
        stores the orbital mutual-information matrix of the solver's state, in the orbital basis it was given, in
        ``self.mutual_info``; a solver that can supply the 3- / 4-body expectation values of every orbital pair does so
        through an optional ``make_pair_supplement()`` -> (n(n-1)/2, 9) method next to make_rdm1 / make_rdm2 (pairs in the
        order i > j, k = i(i-1)/2 + j; the nine numbers are listed in qio_b200/csrc/mi_formulas.inc).  Without it the
        two-orbital RDMs are closed with the 2-body data only (exact for two-electron states).
        """
        self.mf = mf
        self.no = len(mf.mo_coeff)
        self.nocc = int(mf.mol.nelectron // 2)
        self.n_cas, self.n_act_e = act_space
        self.n_core = self.nocc - int(self.n_act_e // 2)
        self.active_indices = np.array(list(range(self.n_core, self.n_core + self.n_cas)))
        self.sol = sol
        self.max_cycle = 100
        self.thresh = 1e-4
        self.level_shift = 1e-3
        self.step_size = 1.
        self.mu = 0.
        self.mu_rate = 1.0
        self.mo_coeff = None
        self.s_val = None
        self.occ_num = None
        self.gamma = None
        self.Gamma = None
        self.verbose = mf.verbose
        self.keep_on_device = False
        self.compute_mutual_info = False
        self.mutual_info = None

        if logger is not None:
            self.logger = logger
        else:
            self.logger = logging.getLogger('qio')
            if not any(type(h) is logging.FileHandler for h in self.logger.handlers):
                self.logger.setLevel(logging.INFO)
                fh = logging.FileHandler(log_file, mode='w')
                fh.setLevel(logging.INFO)
                fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
                self.logger.addHandler(fh)
            if not any(type(h) is logging.StreamHandler for h in self.logger.handlers):
                self.logger.addHandler(logging.StreamHandler())

        self.debug = False

    def dump_flags(self):
        print("=" * 35 + "QICAS FLAGS" + "=" * 35)
        self.logger.info("max_cycle = %d", self.max_cycle)
        self.logger.info("n_core = %d", self.n_core)
        self.logger.info("thresh = %e", self.thresh)
        self.logger.info("level_shift = %e", self.level_shift)
        self.logger.info("step_size = %f", self.step_size)
        self.logger.info("solver = %s", self.sol.__class__.__name__)

    def kernel(self, inactive_indices, mo_coeff, method='newton-raphson'):
        """Run the solver on ``mo_coeff``, optimise the orbitals, return the new MO coefficients.

        method: '2d_jacobi' / '2djacobi' / 'jacobi' or 'newton-raphson' / 'newton_raphson' / 'newton' / 'nr'
        (case-insensitive); anything else raises NotImplementedError (qio/qio.py:134-142)."""
        if mo_coeff is None:
            mo_coeff = self.mf.mo_coeff

        self.sol.kernel(mo_coeff.copy())
        dm1 = self.sol.make_rdm1()
        dm2 = self.sol.make_rdm2()

        n_act_e = np.sum(dm1.diagonal()[self.active_indices])
        self.logger.info("Number of active electrons = %.6f", n_act_e)

        if self.debug:
            nel = self.mf.mol.nelectron
            assert np.isclose(np.sum(dm1.diagonal()), nel, atol=1e-6)
            assert np.isclose(np.einsum('iijj->', dm2), nel * (nel - 1), atol=1e-6)
            occ, _ = np.linalg.eigh(dm1)
            if np.any(occ < 0):
                raise ValueError("Negative eigenvalues in dm1")

        m = method.upper()
        if m not in ('2D_JACOBI', '2DJACOBI', 'JACOBI', 'NEWTON-RAPHSON', 'NEWTON_RAPHSON', 'NEWTON', 'NR'):
            raise NotImplementedError('Only 2d_jacobi and newton-raphson are supported')

        gamma, Gamma = prep_rdm12(dm1, dm2, on_device=True)

        if self.compute_mutual_info:
            from .mutual_info import mutual_information
            make_supp = getattr(self.sol, 'make_pair_supplement', None)
            supp = make_supp() if callable(make_supp) else None
            self.mutual_info = dev.to_host(mutual_information(gamma, Gamma, supp))

        if m in ('2D_JACOBI', '2DJACOBI', 'JACOBI'):
            _, U, g, G = minimize_orb_corr_jacobi(gamma, Gamma, inactive_indices, self.max_cycle)
        else:
            U, g, G, self.mu = minimize_orb_corr_GD(gamma, Gamma, inactive_indices, self.active_indices,
                                                    thresh=self.thresh, max_cycle=self.max_cycle,
                                                    step_size=self.step_size, level_shift=self.level_shift)
        del gamma, Gamma

        V, self.s_val, self.occ_num = reorder_occ(g, G)      # reads 3n diagonal entries from the device
        if self.keep_on_device:
            self.gamma, self.Gamma = g, G
        else:
            self.gamma, self.Gamma = dev.to_host(g), dev.to_host(G)
        U_ = np.matmul(V, U)
        self.mo_coeff = mo_coeff @ U_.T
        return self.mo_coeff


def upload_and_prep_rdm2(dm2, n: int, a0: int = 0, na: int | None = None, slabs_per_chunk: int | None = None, out=None):
    """Host dm2[a0:a0+na] -> device Gamma[a0:a0+na] without ever holding the raw dm2 on the device:
    slabs of the leading index go through two pinned staging buffers on a copy stream while the prep
    kernel of the previous chunk runs (both permutations keep the leading index).
    ``out``: optional preallocated (na, n, n, n) device tensor that receives Gamma."""
    d = dev.require_cuda()
    na = n - a0 if na is None else na
    Gamma = dev.empty(na, n, n, n) if out is None else out
    if na == 0:
        return Gamma
    slab = n ** 3
    if slabs_per_chunk is None:
        slabs_per_chunk = max(1, min(na, (64 << 20) // (8 * slab) or 1))
    compute = torch.cuda.current_stream()
    copy = torch.cuda.Stream()
    staged = [torch.empty((slabs_per_chunk, slab), dtype=torch.float64, device=d) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_pinned = isinstance(dm2, torch.Tensor) and dm2.is_pinned() and dm2.is_contiguous() and dm2.dtype == torch.float64
    if host_pinned:
        src_t = dm2[a0:a0 + na].reshape(na, slab)      # caller-pinned memory: DMA straight out of it
    else:
        src = np.ascontiguousarray(np.asarray(dm2)[a0:a0 + na], dtype=np.float64).reshape(na, slab)
        pinned = [torch.empty((slabs_per_chunk, slab), dtype=torch.float64).pin_memory() for _ in range(2)]
        host_free = [torch.cuda.Event() for _ in range(2)]
    copy.wait_stream(compute)
    for c, s0 in enumerate(range(0, na, slabs_per_chunk)):
        k = min(slabs_per_chunk, na - s0)
        b = c & 1
        if not host_pinned:
            if c >= 2:
                host_free[b].synchronize()           # pinned buffer b has been read by its H2D copy
            pinned[b][:k].numpy()[...] = src[s0:s0 + k]
        with torch.cuda.stream(copy):
            if c >= 2:
                copy.wait_event(consumed[b])     # device staging buffer b has been consumed by its prep kernel
            staged[b][:k].copy_(src_t[s0:s0 + k] if host_pinned else pinned[b][:k], non_blocking=True)
            if not host_pinned:
                host_free[b].record(copy)
            copied[b].record(copy)
        compute.wait_event(copied[b])
        dev.prep_rdm12(None, staged[b][:k], n, k, Gamma=Gamma[s0:s0 + k])
        consumed[b].record(compute)
    compute.synchronize()
    return Gamma


def prep_rdm12(dm1, dm2, on_device=False):
    """Spin-interleaved 1-RDM and the (up, down, down, up) block of the 2-RDM from pyscf-convention
    spatial RDMs of a singlet (qio/qio.py:152-174):
        rdm1[::2, ::2] = rdm1[1::2, 1::2] = dm1 / 2
        rdm2[a,b,c,d] = (2 dm2[a,d,b,c] + dm2[a,c,b,d]) / 6
    numpy in -> numpy out (unless ``on_device``); CUDA tensors in -> CUDA tensors out."""
    n = len(dm1)
    d1 = dev.to_device(dm1)
    if dev.is_device(dm2):
        gamma, Gamma = dev.prep_rdm12(d1, dm2.contiguous(), n)
        return gamma, Gamma
    gamma, _ = dev.prep_rdm12(d1, None, n, na=0, Gamma=dev.empty(0))
    Gamma = upload_and_prep_rdm2(dm2, n)
    if on_device:
        return gamma, Gamma
    return dev.to_host(gamma), dev.to_host(Gamma)


def get_overlap(fcivector, cisdvec):
    """Normalised overlap of two CI vectors (qio/qio.py:177-189); host-side helper outside the hot path."""
    return np.dot(fcivector, cisdvec) / (np.linalg.norm(fcivector) * np.linalg.norm(cisdvec))
