"""Solver-side helpers (reference qio/solver/): the pieces of the TCCSD wrapper that are dense FP64 linear algebra and do not
need pyscf -- the CAS <-> CCSD amplitude projections (tccsd.py:153-172) on the device, natural orbitals and
semi-canonicalisation (tccsd.py:209-265) on the host.  The wrappers themselves (pyscf CCSD / CASCI, block2 DMRG) stay
the reference's own and plug into qio_b200.QIO unchanged."""
from . import tccsd  # noqa: F401
