"""Optimiser layer (reference qio/grad/): Jacobi sweeps and the diagonal-Newton driver."""
from . import gradient, jacobi  # noqa: F401
