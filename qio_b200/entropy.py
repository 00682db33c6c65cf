"""Entropy layer -- same functions and signatures as the reference's ``qio/entropy.py``
(shannon :4-22, get_cost_fqi :24-47), evaluated by the CUDA library.

``gamma`` / ``Gamma`` may be numpy arrays (uploaded for the call) or CUDA tensors (used in place;
keep them on the device between calls, e.g. ``QIO.keep_on_device = True``).
"""
from __future__ import annotations

from warnings import warn

import numpy as np

from . import device as dev
from ._lib import FLAG_RAISE_GT1, FLAG_WARN_NEGATIVE


def raise_or_warn(flags: int, spec=None) -> None:
    """Host side of the reference's checks (qio/entropy.py:15-20) from the device flags."""
    if flags & FLAG_WARN_NEGATIVE:
        warn("Warning: spec has negative entries!")
    if flags & FLAG_RAISE_GT1:
        if spec is not None:
            print(spec)
        raise ValueError("spec has entries larger than 1")


def shannon(spec):
    """Shannon entropy -sum_{p>0} p ln p of a probability distribution of any shape.
    Warns on entries below -1e-6, raises ValueError on entries above 1 when none is negative."""
    p = dev.to_device(spec).reshape(-1)
    total, _, _, flags = dev.to_host(dev.shannon(p))
    if int(flags):
        raise_or_warn(int(flags), dev.to_host(p).reshape(np.shape(spec)))
    return np.float64(total)


def get_cost_fqi(gamma, Gamma, inactive_indices):
    """Sum of the one-orbital entropies of the inactive orbitals (qio/entropy.py:24-47)."""
    n = int(Gamma.shape[0])
    inds = dev.to_device(np.asarray(inactive_indices, dtype=np.int32), dev.I32)
    g = dev.to_device(gamma)
    G = dev.to_device(Gamma)
    diag = dev.orbital_diagonals(g, G, n, inds)
    spec, _, _, summary = dev.orbital_entropies(diag, want_spec=True)
    total, _, _, flags = dev.to_host(summary)
    if int(flags):
        raise_or_warn(int(flags), dev.to_host(spec))
    return np.float64(total)
