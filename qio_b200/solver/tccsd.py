"""Dense linear algebra of the reference's tailored-CCSD wrapper (qio/solver/tccsd.py) without its pyscf parts.

    project_t1 / project_t2          the amplitude projections between the CAS and the CCSD orbital spaces
                                     einsum('ia,Ii,Aa->IA'), einsum('ijab,Ii,Jj,Aa,Bb->IJAB') and their reverses
                                     (tccsd.py:153-154,157-158,164-172)
    make_tailor_callback             the body of ``callback`` (tccsd.py:160-175) as a function of the FCI amplitudes and the
                                     two overlap matrices -- what ``cc.callback`` is set to
    make_no                          natural orbitals of a (sub)space (tccsd.py:209-244)
    semi_canonicalize                occupied / virtual Fock diagonalisation (tccsd.py:247-267)

The T2 projection is the same computation as the 4-index rotation of the hot path -- four single-axis FP64 contractions,
here with rectangular matrices -- and runs on the rotation kernels (rotate_tma.cu / rotate.cu, FP64 DMMA); T1 and the two
n x n eigenproblems are O(n^3) host numpy.  A maintainer replaces the four ``einsum`` lines of the reference's ``callback``
by ``project_t1`` / ``project_t2``; everything that needs pyscf stays where it is.
"""
from __future__ import annotations

from functools import reduce

import numpy as np

from .. import device as dev


def project_t1(t1, pocc, pvir, to_cas=False):
    """t1 (o, v) of the CAS -> (O, V) of the CCSD space: einsum('ia,Ii,Aa->IA', t1, pocc, pvir) (tccsd.py:153,170);
    ``to_cas``: the reverse einsum('IA,Ii,Aa->ia') (tccsd.py:164).  pocc (O, o), pvir (V, v).  Host numpy (O(n^3))."""
    t1, pocc, pvir = np.asarray(t1), np.asarray(pocc), np.asarray(pvir)
    if to_cas:
        return reduce(np.dot, (pocc.T, t1, pvir))
    return reduce(np.dot, (pocc, t1, pvir.T))


def project_t2(t2, pocc, pvir, to_cas=False):
    """t2 (o, o, v, v) of the CAS -> (O, O, V, V) of the CCSD space: einsum('ijab,Ii,Jj,Aa,Bb->IJAB') (tccsd.py:154,171);
    ``to_cas``: the reverse einsum('IJAB,Ii,Jj,Aa,Bb->ijab') (tccsd.py:165).  Four cyclic single-axis contractions on the
    device; numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    Po = dev.to_device(np.ascontiguousarray(np.asarray(dev.to_host(pocc) if dev.is_device(pocc) else pocc).T) if to_cas else pocc)
    Pv = dev.to_device(np.ascontiguousarray(np.asarray(dev.to_host(pvir) if dev.is_device(pvir) else pvir).T) if to_cas else pvir)
    src = dev.to_device(t2).contiguous()
    shape = list(src.shape)
    for M in (Po, Po, Pv, Pv):                          # (i,j,a,b) -> (j,a,b,I) -> (a,b,I,J) -> (b,I,J,A) -> (I,J,A,B)
        n_out, kdim = int(M.shape[0]), int(M.shape[1])
        if kdim != shape[0]:
            raise ValueError(f'project_t2: projection matrix {tuple(M.shape)} does not match the amplitude axis of length {shape[0]}')
        rows = int(np.prod(shape[1:]))
        dst = dev.empty(rows, n_out)
        dev.contract_axis_cyclic(M.contiguous(), src, dst, n_out, kdim, rows)
        shape = shape[1:] + [n_out]
        src = dst
    out = src.view(*shape)
    return dev.like_input(out, t2)


def make_tailor_callback(t1cas_fci, t2cas_fci, pocc, pvir):
    """``callback(kwargs)`` of tccsd.py:160-175: project the CCSD amplitudes onto the CAS, take the difference to the FCI
    amplitudes there, rotate it back and add it -- in place on kwargs['t1new'], kwargs['t2new'] (numpy arrays)."""
    t1cas_fci, t2cas_fci = np.asarray(t1cas_fci), np.asarray(t2cas_fci)
    Po_d, Pv_d = dev.to_device(pocc), dev.to_device(pvir)
    t2fci_d = dev.to_device(t2cas_fci)

    def callback(kwargs):
        t1, t2 = kwargs['t1new'], kwargs['t2new']
        t1cas_cc = project_t1(t1, pocc, pvir, to_cas=True)
        t2cas_cc = project_t2(dev.to_device(t2), Po_d, Pv_d, to_cas=True)
        dt1 = project_t1(t1cas_fci - t1cas_cc, pocc, pvir)
        dt2 = project_t2(t2fci_d - t2cas_cc, Po_d, Pv_d)
        t1 += dt1
        t2 += dev.to_host(dt2)

    return callback


def make_no(rdm1, mo_coeff, subspace=None):
    """Natural orbitals from a 1-RDM, optionally within ``subspace`` (tccsd.py:209-244).  Returns (no_coeff, rdm1_new).

    As in the reference, with a subspace ``rdm1_new`` is an unchanged copy of ``rdm1``: the reference assigns the diagonal
    2 into ``rdm1_new[subspace,:][:,subspace]``, which is a temporary (tccsd.py:238)."""
    rdm1, mo_coeff = np.asarray(rdm1), np.asarray(mo_coeff)
    if subspace is not None:
        rdm1_sub = rdm1[subspace, :][:, subspace]
        mo_coeff_sub = mo_coeff[:, subspace]
    else:
        rdm1_sub, mo_coeff_sub = rdm1, mo_coeff
    e, v = np.linalg.eigh(rdm1_sub)
    idx = np.argsort(e)[::-1]
    v = v[:, idx]
    no_coeff_sub = np.dot(mo_coeff_sub, v)
    no_coeff = mo_coeff.copy()
    if subspace is not None:
        no_coeff[:, subspace] = no_coeff_sub
        rdm1_new = rdm1.copy()
    else:
        no_coeff = no_coeff_sub
        rdm1_new = np.diag(np.ones(len(rdm1)) * 2)
    return no_coeff, rdm1_new


def semi_canonicalize(mf, mo_coeff, nocc):
    """Diagonalise the occupied-occupied and the virtual-virtual Fock blocks (tccsd.py:247-267); uses ``mf.get_fock()``.
    Returns the semi-canonical MO coefficients."""
    fockao = mf.get_fock()
    fockmo = reduce(np.dot, (mo_coeff.conj().T, fockao, mo_coeff))
    _, v_canon_occ = np.linalg.eigh(fockmo[:nocc, :nocc])
    _, v_canon_vir = np.linalg.eigh(fockmo[nocc:, nocc:])
    return np.concatenate((np.dot(mo_coeff[:, :nocc], v_canon_occ), np.dot(mo_coeff[:, nocc:], v_canon_vir)), axis=1)
