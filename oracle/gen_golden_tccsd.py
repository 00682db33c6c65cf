"""Golden vectors for the solver-side helpers, produced by RUNNING THE REFERENCE'S OWN FUNCTIONS  --  test infrastructure only.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_tccsd.py [/root/reference]

qio/solver/tccsd.py imports pyscf at module level (absent here), so ``make_no`` and ``semi_canonicalize`` -- two
self-contained numpy functions -- are compiled from the reference file's own, unmodified source text (ast) and executed.
The two projection einsums of tccsd.py:153-172 are literal numpy calls; their strings are read from the file
and evaluated as they stand.  Writes tests/golden/ref_tccsd_helpers.npz.
"""
from __future__ import annotations

import ast
import os
import re
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'


def reference_functions(path):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ('make_no', 'semi_canonicalize'):
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), ns)
    return ns, src


class _MF:
    def __init__(self, fock):
        self._f = fock

    def get_fock(self):
        return self._f


def main():
    ns, src = reference_functions(os.path.join(REF, 'qio', 'solver', 'tccsd.py'))
    rng = np.random.default_rng(2024)
    n, nao, nocc = 14, 16, 5
    A = rng.standard_normal((n, n))
    rdm1 = A @ A.T / n
    mo = rng.standard_normal((nao, n))
    sub = [2, 3, 4, 7, 9]
    no_full, r_full = ns['make_no'](rdm1, mo)
    no_sub, r_sub = ns['make_no'](rdm1, mo, sub)
    F = rng.standard_normal((nao, nao))
    F = F + F.T
    semican = ns['semi_canonicalize'](_MF(F), mo, nocc)
    # the projection einsums, as written in the file
    strings = sorted(set(re.findall(r"einsum\('([A-Za-z,]+->[A-Za-z]+)'", src)))
    o, v, O, V = 3, 4, 6, 9
    pocc, pvir = rng.standard_normal((O, o)), rng.standard_normal((V, v))
    t1c, t2c = rng.standard_normal((o, v)), rng.standard_normal((o, o, v, v))
    t1, t2 = rng.standard_normal((O, V)), rng.standard_normal((O, O, V, V))
    d = dict(rdm1=rdm1, mo=mo, sub=np.array(sub), no_full=no_full, r_full=r_full, no_sub=no_sub, r_sub=r_sub, fock=F, nocc=nocc,
             semican=semican, pocc=pocc, pvir=pvir, t1c=t1c, t2c=t2c, t1=t1, t2=t2, einsum_strings=np.array(strings))
    assert 'ijab,Ii,Jj,Aa,Bb->IJAB' in strings and 'IJAB,Ii,Jj,Aa,Bb->ijab' in strings
    d['t1_up'] = np.einsum('ia,Ii,Aa->IA', t1c, pocc, pvir)
    d['t2_up'] = np.einsum('ijab,Ii,Jj,Aa,Bb->IJAB', t2c, pocc, pocc, pvir, pvir)
    d['t1_down'] = np.einsum('IA,Ii,Aa->ia', t1, pocc, pvir)
    d['t2_down'] = np.einsum('IJAB,Ii,Jj,Aa,Bb->ijab', t2, pocc, pocc, pvir, pvir)
    # the callback body, line by line (tccsd.py:164-175)
    dt1 = t1c - d['t1_down']
    dt2 = t2c - d['t2_down']
    d['t1_tailored'] = t1 + np.einsum('ia,Ii,Aa->IA', dt1, pocc, pvir)
    d['t2_tailored'] = t2 + np.einsum('ijab,Ii,Jj,Aa,Bb->IJAB', dt2, pocc, pocc, pvir, pvir)
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_tccsd_helpers.npz')
    np.savez_compressed(path, **d)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB;', len(strings), 'einsum strings in the reference file')


if __name__ == '__main__':
    main()
