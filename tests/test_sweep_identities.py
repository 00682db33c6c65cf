"""CPU checks of the two algebraic identities the pipelined sweep kernel (qio_b200/csrc/sweep3.cu) rests on, against the
oracle's dense Givens transform (restatement of qio/grad/jacobi.py:63-115).  No GPU, no product code: documentation that runs."""
import numpy as np

from oracle import qio_oracle as O


def _random_state(n, seed):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, n, n, n))
    g = rng.standard_normal((2 * n, 2 * n))
    return g, G


def test_block_of_pair_k_is_a_linear_image_of_a_three_pair_super_block():
    """B_k = T (x4) P_k: the 16-number block of pair k after rotations k-2 and k-1 equals the contraction of the super block
    over the orbitals D of pairs k-2, k-1, k (taken BEFORE those rotations) with T = rows of G_{k-1} G_{k-2} of the two
    target orbitals (sweep3.cu: search_main)."""
    n = 7
    g, G = _random_state(n, 0)
    pairs = [(4, 1), (4, 6), (2, 6)]                      # pairs k-2, k-1, k (k-1 and k do not even share an orbital)
    thetas = [0.37, -1.12]
    D = []
    for p in pairs:
        for x in p:
            if x not in D:
                D.append(x)
    m = len(D)
    P = G[np.ix_(D, D, D, D)]                             # super block of the state before rotation k-2
    # the kernel's T: start from unit rows, apply the pending rotations oldest first (rows i, j of the coefficient matrix)
    C = np.eye(m)
    for (i, j), t in zip(pairs[:2], thetas):
        a, b = D.index(i), D.index(j)
        c, s = np.cos(t), np.sin(t)
        ra, rb = C[a].copy(), C[b].copy()
        C[a], C[b] = c * ra + s * rb, c * rb - s * ra
    ik, jk = pairs[2]
    T = C[[D.index(ik), D.index(jk)]]
    B = np.einsum('ap,bq,cr,ds,pqrs->abcd', T, T, T, T, P)
    # reference: apply the two rotations to the full tensor, then read the block of pair k
    G2 = G
    g2 = g
    for (i, j), t in zip(pairs[:2], thetas):
        g2, G2 = O.jacobi_transform(g2, G2, i, j, t)
    ref = G2[np.ix_([ik, jk], [ik, jk], [ik, jk], [ik, jk])]
    assert np.allclose(B, ref, atol=1e-13, rtol=0)


def test_rows_of_a_dormant_orbital_are_caught_up_by_the_accumulated_rotation():
    """Skipping the rows (y, x) and (x, y) of an orbital x that takes part in no rotation, and applying the accumulated
    rotation R to their other index afterwards, gives the same tensor as rotating everything (sweep3.cu: catch_up).  Axes 1
    and 2 only -- the kernel defers axes 0 and 3 -- so the comparison is on those two axes."""
    n, x = 6, 3
    rng = np.random.default_rng(1)
    S = rng.standard_normal((n, n, n, n))
    rots = [((0, 1), 0.4), ((0, 4), -0.9), ((5, 1), 1.3), ((4, 5), 0.2)]      # none involves x
    full = S.copy()
    R = np.eye(n)
    for (i, j), t in rots:
        V = O.givens_matrix(n, i, j, t)
        full = np.einsum('bB,cC,aBCd->abcd', V, V, full)
        R = V @ R
    # lazy: rotate only the rows whose free index is not x, then catch up
    lazy = S.copy()
    keep = [y for y in range(n) if y != x]
    for (i, j), t in rots:
        V = O.givens_matrix(n, i, j, t)
        sub = lazy[np.ix_(range(n), keep, keep, range(n))]
        lazy[np.ix_(range(n), keep, keep, range(n))] = np.einsum('bB,cC,aBCd->abcd', V[np.ix_(keep, keep)], V[np.ix_(keep, keep)], sub)
    A = keep
    lazy[:, A, x, :] = np.einsum('yY,aYd->ayd', R[np.ix_(A, A)], lazy[:, A, x, :])       # column x: R over the first index
    lazy[:, x, A, :] = np.einsum('yY,aYd->ayd', R[np.ix_(A, A)], lazy[:, x, A, :])       # row x: R over the second index
    assert np.allclose(lazy, full, atol=1e-13, rtol=0)
