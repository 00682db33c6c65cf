"""Oracle for the two-orbital RDM / pair entropy / mutual information EXTENSION.

TEST INFRASTRUCTURE ONLY, and **parity unpinned**: the reference (schilling-group/QIO) has no two-orbital
RDM, no eigensolve and no mutual information anywhere, so nothing here can be checked against it.  What pins
this file instead is brute force: explicit many-electron wavefunctions in Fock space, whose two-orbital
reduced density matrices are obtained by a literal partial trace.

Definitions.  Orbital k has the local states |0>, |up>, |dn>, |updn> = a_up^+ a_dn^+ |0>.  For a pair (i, j) the
16 states |s_i s_j> are built with the creation operators in the order i_up, i_dn, j_up, j_dn.  rho_ij[s, s'] =
<psi| (|s'><s|) |psi>.  It is block diagonal in (N, S_z): block sizes 1, 2, 2, 1, 4, 1, 2, 2, 1 (``BLOCKS``).
Every |s'><s| is a polynomial in the four modes' creation / annihilation operators; expanding it in normal-ordered
strings  c^+_A c_B  gives rho_ij as a LINEAR function of k-body expectation values:
  k = 0: 1;  k = 1: gamma;  k = 2: the 2-RDM (opposite spin = Gamma, same spin = Gamma[abcd] - Gamma[abdc] for a
  singlet);  k = 3, 4: not contained in rdm1 / rdm2 -> the "supplement" (zeros by default; exact for two electrons).
The expansion coefficients are found numerically (16 x 16 Jordan-Wigner matrices, a 256 x 256 linear solve), so no
sign is derived by hand; ``FORMULAS`` is the table, and ``oracle/gen_mi_formulas.py`` prints it as the C include
used by the CUDA kernel.
"""
from __future__ import annotations

import itertools

import numpy as np

MODES = ('iu', 'id', 'ju', 'jd')           # mode order inside the pair: i_up, i_dn, j_up, j_dn


# ---------------------------------------------------------------------------------------------------------
# Fock-space tools (Jordan-Wigner), generic number of modes
# ---------------------------------------------------------------------------------------------------------
def annihilators(m):
    """List of m matrices (2^m x 2^m) c_0..c_{m-1}; basis index bit k (LSB = mode 0) = occupation of mode k,
    |occ> = prod_k (c_k^+)^{occ_k} |0> with the product taken in increasing mode order (mode 0 leftmost)."""
    dim = 1 << m
    ops = []
    for k in range(m):
        c = np.zeros((dim, dim))
        for s in range(dim):
            if (s >> k) & 1:
                sign = (-1) ** bin(s & ((1 << k) - 1)).count('1')      # modes to the left of k
                c[s & ~(1 << k), s] = sign
        ops.append(c)
    return ops


def string_matrix(cre, ann, ops):
    """c^+_{cre[0]} c^+_{cre[1]} ... c_{ann[0]} c_{ann[1]} ... as a matrix."""
    dim = ops[0].shape[0]
    M = np.eye(dim)
    for k in cre:
        M = M @ ops[k].T
    for k in ann:
        M = M @ ops[k]
    return M


# local pair Fock space: 4 modes, 16 states.  State label (s_i, s_j) with s in {0: empty, 1: up, 2: dn, 3: updn}
_OPS4 = annihilators(4)


def pair_state_index(si, sj):
    """Fock index (bit k = occupation of MODES[k]) of |s_i s_j>; signs are those of the ordered product."""
    occ = [(si & 1), (si >> 1) & 1, (sj & 1), (sj >> 1) & 1]      # s = 1 -> up only, 2 -> dn only, 3 -> both
    return occ[0] | (occ[1] << 1) | (occ[2] << 2) | (occ[3] << 3)


def _sector(idx):
    occ = [(idx >> k) & 1 for k in range(4)]
    return sum(occ), occ[0] + occ[2] - occ[1] - occ[3]          # (N, 2 S_z)


# blocks in the order N = 0; 1 up; 1 dn; 2 upup; 2 S_z=0; 2 dndn; 3 up; 3 dn; 4
_SECTOR_ORDER = [(0, 0), (1, 1), (1, -1), (2, 2), (2, 0), (2, -2), (3, 1), (3, -1), (4, 0)]
BLOCKS = [[idx for idx in range(16) if _sector(idx) == sec] for sec in _SECTOR_ORDER]
BLOCK_SIZES = [len(b) for b in BLOCKS]
assert BLOCK_SIZES == [1, 2, 2, 1, 4, 1, 2, 2, 1]


# ---------------------------------------------------------------------------------------------------------
# Expansion of the matrix units in normal-ordered strings
# ---------------------------------------------------------------------------------------------------------
def _all_strings():
    out = []
    for ka in range(5):
        for A in itertools.combinations(range(4), ka):
            for kb in range(5):
                for B in itertools.combinations(range(4), kb):
                    out.append((A, B))
    return out


def _spin(mode):
    return +1 if mode in (0, 2) else -1


def _orb(mode):
    return 0 if mode < 2 else 1          # 0 = i, 1 = j


def source_of_string(A, B):
    """Map <c^+_A c_B> (A, B increasing mode tuples, |A| = |B| = k) to (kind, index, sign):
    kind 'const' | 'rec' (index into the 24-double pair block) | 'same' (same-spin 2-RDM entry, evaluated as
    rec[t1] - rec[t2]) | 'supp' (index into the supplement list SUPP_STRINGS)."""
    k = len(A)
    if k == 0:
        return ('const', 0, 1.0)
    if k == 1:
        a, b = A[0], B[0]
        if _spin(a) != _spin(b):
            return None                                    # spin non-conserving: zero in an S_z eigenstate
        blk = 16 if _spin(a) > 0 else 20
        return ('rec', blk + _orb(a) * 2 + _orb(b), 1.0)   # gamma[2I_m+s, 2I_n+s], m = orb(a), n = orb(b)
    if k == 2:
        (a1, a2), (b1, b2) = A, B
        if _spin(a1) + _spin(a2) != _spin(b1) + _spin(b2):
            return None
        if _spin(a1) != _spin(a2):
            # creators: one up, one dn.  Canonical element <a_up^+ b_dn^+ c_dn d_up> = Gamma[a, b, c, d]
            up_c, dn_c = (a1, a2) if _spin(a1) > 0 else (a2, a1)
            s = 1.0 if _spin(a1) > 0 else -1.0             # c^+_{dn} c^+_{up} = - c^+_{up} c^+_{dn}
            # annihilators in canonical order: c_dn then c_up
            dn_a, up_a = (b1, b2) if _spin(b1) < 0 else (b2, b1)
            s *= 1.0 if _spin(b1) < 0 else -1.0
            t = _orb(up_c) * 8 + _orb(dn_c) * 4 + _orb(dn_a) * 2 + _orb(up_a)
            return ('rec', t, s)
        # same spin: <a_s^+ b_s^+ c_s d_s> = Gamma[a,b,c,d] - Gamma[a,b,d,c]   (singlet)
        t1 = _orb(a1) * 8 + _orb(a2) * 4 + _orb(b1) * 2 + _orb(b2)
        t2 = _orb(a1) * 8 + _orb(a2) * 4 + _orb(b2) * 2 + _orb(b1)
        return ('same', (t1, t2), 1.0)
    return ('supp', SUPP_INDEX[(A, B)], 1.0)


# canonical list of the k >= 3 strings that conserve N and S_z
SUPP_STRINGS = [(A, B) for (A, B) in _all_strings()
                if len(A) == len(B) >= 3 and sum(map(_spin, A)) == sum(map(_spin, B))]
SUPP_INDEX = {ab: k for k, ab in enumerate(SUPP_STRINGS)}
N_SUPP = len(SUPP_STRINGS)


def _build_formulas():
    strings = _all_strings()
    basis = np.stack([string_matrix(A, B, _OPS4).ravel() for (A, B) in strings], axis=1)      # 256 x 256
    formulas = []        # per block: matrix of term lists
    for blk in BLOCKS:
        rows = []
        for s in blk:
            row = []
            for sp in blk:
                unit = np.zeros((16, 16))
                unit[sp, s] = 1.0                           # |s'><s|
                coef = np.linalg.solve(basis, unit.ravel())
                terms = []
                for c, (A, B) in zip(coef, strings):
                    if abs(c) < 1e-12:
                        continue
                    assert abs(c - round(c)) < 1e-9
                    if len(A) != len(B):
                        continue                            # particle-number non-conserving: zero
                    src = source_of_string(A, B)
                    if src is None:
                        continue
                    kind, idx, sign = src
                    terms.append((float(round(c)) * sign, kind, idx))
                row.append(terms)
            rows.append(row)
        formulas.append(rows)
    return formulas


FORMULAS = _build_formulas()


def two_orbital_rdm_from_rdms(block24, supp=None):
    """List of the 9 (N, S_z) blocks of rho_ij from the pair block (16 Gamma + 4 gamma_up + 4 gamma_dn entries)
    and the optional supplement (N_SUPP numbers, zeros by default)."""
    rec = np.asarray(block24, dtype=np.float64)
    sp = np.zeros(N_SUPP) if supp is None else np.asarray(supp, dtype=np.float64)
    out = []
    for rows in FORMULAS:
        m = len(rows)
        M = np.zeros((m, m))
        for r in range(m):
            for c in range(m):
                v = 0.0
                for coef, kind, idx in rows[r][c]:
                    if kind == 'const':
                        v += coef
                    elif kind == 'rec':
                        v += coef * rec[idx]
                    elif kind == 'same':
                        v += coef * (rec[idx[0]] - rec[idx[1]])
                    else:
                        v += coef * sp[idx]
                M[r, c] = v
        out.append(M)
    return out


def entropy_of_eigenvalues(lam):
    lam = np.asarray(lam)
    lam = lam[lam > 0]
    return float(-np.sum(lam * np.log(lam)))


def pair_entropy(block24, supp=None):
    """(S_ij, sorted-by-block eigenvalues (16,)) of the two-orbital RDM; blocks are symmetrised before eigvalsh."""
    eig = []
    for M in two_orbital_rdm_from_rdms(block24, supp):
        eig.extend(np.linalg.eigvalsh(0.5 * (M + M.T)).tolist())
    eig = np.array(eig)
    return entropy_of_eigenvalues(eig), eig


# ---------------------------------------------------------------------------------------------------------
# Brute force on explicit wavefunctions (n orbitals -> 2n modes, mode 2k = k_up, 2k+1 = k_dn)
# ---------------------------------------------------------------------------------------------------------
class FockState:
    def __init__(self, psi, n):
        self.n = n
        self.psi = np.asarray(psi, dtype=np.float64)
        self.psi = self.psi / np.linalg.norm(self.psi)
        self.ops = annihilators(2 * n)

    def expect(self, cre, ann):
        """<c^+_{cre...} c_{ann...}> with global mode numbers."""
        return float(self.psi @ (string_matrix(cre, ann, self.ops) @ self.psi))

    def rdms(self):
        """gamma (2n x 2n, interleaved spin) and Gamma[a,b,c,d] = <a_up^+ b_dn^+ c_dn d_up> (qio/qio.py conventions)."""
        n = self.n
        gamma = np.zeros((2 * n, 2 * n))
        for p in range(2 * n):
            for q in range(2 * n):
                gamma[p, q] = self.expect([p], [q])
        Gamma = np.zeros((n, n, n, n))
        for a, b, c, d in itertools.product(range(n), repeat=4):
            Gamma[a, b, c, d] = self.expect([2 * a, 2 * b + 1], [2 * c + 1, 2 * d])
        return gamma, Gamma

    def pair_modes(self, i, j):
        return [2 * i, 2 * i + 1, 2 * j, 2 * j + 1]

    def supplement(self, i, j):
        g = self.pair_modes(i, j)
        return np.array([self.expect([g[a] for a in A], [g[b] for b in B]) for (A, B) in SUPP_STRINGS])

    def two_orbital_rdm(self, i, j):
        """16 x 16 rho_ij[s, s'] = <psi| |s'><s| |psi> with the local operators embedded in the full Fock space."""
        g = self.pair_modes(i, j)
        strings = _all_strings()
        basis = np.stack([string_matrix(A, B, _OPS4).ravel() for (A, B) in strings], axis=1)
        vals = np.array([self.expect([g[a] for a in A], [g[b] for b in B]) for (A, B) in strings])
        rho = np.zeros((16, 16))
        for s in range(16):
            for sp in range(16):
                unit = np.zeros((16, 16))
                unit[sp, s] = 1.0
                rho[s, sp] = np.linalg.solve(basis, unit.ravel()) @ vals
        return rho


def random_singlet_state(n, nelec, seed=0):
    """Random vector in the (N = nelec, S_z = 0, S^2 = 0) sector of n orbitals (2n modes)."""
    rng = np.random.default_rng(seed)
    m = 2 * n
    ops = annihilators(m)
    dim = 1 << m
    idx = [s for s in range(dim)
           if bin(s).count('1') == nelec and sum(((s >> k) & 1) * (1 if k % 2 == 0 else -1) for k in range(m)) == 0]
    # S^2 = S_- S_+ + S_z (S_z + 1), S_+ = sum_k c^+_{k up} c_{k dn}
    Sp = sum(ops[2 * k].T @ ops[2 * k + 1] for k in range(n))
    S2 = (Sp.T @ Sp)[np.ix_(idx, idx)]
    w, v = np.linalg.eigh(S2)
    null = v[:, np.abs(w) < 1e-9]
    coef = null @ rng.standard_normal(null.shape[1])
    psi = np.zeros(dim)
    psi[idx] = coef
    return FockState(psi, n)


def geminal_state(C):
    """Two-electron singlet psi = sum_ab C_ab a_up^+ b_dn^+ |0> with symmetric C."""
    n = C.shape[0]
    psi = np.zeros(1 << (2 * n))
    for a in range(n):
        for b in range(n):
            up, dn = 2 * a, 2 * b + 1
            idx = (1 << up) | (1 << dn)
            sign = 1.0 if up < dn else -1.0          # ordered product: lower mode first
            psi[idx] += sign * C[a, b]
    return FockState(psi, n)


def block24_of(gamma, Gamma, i, j):
    I = (i, j)
    rec = np.zeros(24)
    for t in range(16):
        m, n_, p, q = (t >> 3) & 1, (t >> 2) & 1, (t >> 1) & 1, t & 1
        rec[t] = Gamma[I[m], I[n_], I[p], I[q]]
    for u in range(4):
        rec[16 + u] = gamma[2 * I[u >> 1], 2 * I[u & 1]]
        rec[20 + u] = gamma[2 * I[u >> 1] + 1, 2 * I[u & 1] + 1]
    return rec


def mutual_information(gamma, Gamma, supp_fn=None):
    """(n, n) matrix I_ij = S_i + S_j - S_ij (zero diagonal) from rdm1 / rdm2 (+ optional supplement callback)."""
    n = Gamma.shape[0]
    k = np.arange(n)
    nu, nd, nn = gamma[2 * k, 2 * k], gamma[2 * k + 1, 2 * k + 1], Gamma[k, k, k, k]
    spec = np.array([1 - nu - nd + nn, nu - nn, nd - nn, nn])
    S1 = np.array([entropy_of_eigenvalues(spec[:, q]) for q in range(n)])
    I = np.zeros((n, n))
    for i in range(n):
        for j in range(i):
            supp = None if supp_fn is None else supp_fn(i, j)
            Sij, _ = pair_entropy(block24_of(gamma, Gamma, i, j), supp)
            I[i, j] = I[j, i] = S1[i] + S1[j] - Sij
    return I
