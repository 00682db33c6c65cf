"""Recipe for oracle/_ref: an UNMODIFIED copy of the reference's own Python package, taken from the read-only tree
where it lies (/root/reference/qio), so that the benchmark's CPU arm and the tests can run the REAL reference on the
GPU box, where /root/reference does not exist.

    python oracle/make_ref.py [/root/reference]

oracle/_ref/ is listed in .gitignore (the reference's sources never enter this repository's history) but not in
.gpurunignore, so the copy travels with the snapshot like the built .so files.  The reference is pure Python
(numpy + scipy): nothing is compiled, nothing is patched; its solver wrappers (qio/solver/*, need pyscf / block2) are
copied too but never imported.  __graft_entry__.build() calls this when the reference tree is present.
Test / benchmark infrastructure only -- nothing under qio_b200/ may import it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref')


def make_ref(reference_root: str = '/root/reference') -> bool:
    src = os.path.join(reference_root, 'qio')
    if not os.path.isdir(src):
        return False
    tmp = DEST + '.tmp'
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    shutil.copytree(src, os.path.join(tmp, 'qio'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for extra in ('LICENSE', 'README.md'):
        p = os.path.join(reference_root, extra)
        if os.path.exists(p):
            shutil.copy(p, os.path.join(tmp, extra))
    with open(os.path.join(tmp, 'ORIGIN'), 'w') as f:
        f.write(f'unmodified copy of {src} (schilling-group/QIO), made by oracle/make_ref.py\n')
    shutil.rmtree(DEST, ignore_errors=True)
    os.rename(tmp, DEST)
    return True


def load_reference():
    """The reference's hot-path modules (qio.entropy, qio.grad.jacobi, qio.grad.gradient, qio.qio) imported from
    oracle/_ref, or None when the copy does not exist.  The package is loaded under its own name ``qio``; callers must
    not have installed qio_b200 as ``qio`` in the same process."""
    if not os.path.isdir(os.path.join(DEST, 'qio')):
        return None
    import importlib
    import logging
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    mods = {}
    for name in ('qio.entropy', 'qio.grad.jacobi', 'qio.grad.gradient', 'qio.qio'):
        mods[name.split('.')[-1]] = importlib.import_module(name)
    if not os.path.abspath(mods['jacobi'].__file__).startswith(os.path.abspath(DEST)):
        return None         # some other package named qio is shadowing the copy
    logging.getLogger('qio').disabled = True       # the reference has malformed log calls
    return mods


if __name__ == '__main__':
    ok = make_ref(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('oracle/_ref', 'written' if ok else 'NOT written (reference tree not found)')
