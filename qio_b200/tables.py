"""Angle tables of the line search, built with numpy ON THE HOST.

The reference scans ``np.arange(0.01, pi, 0.01)`` and then ``np.arange(t_opt-0.01, t_opt+0.01, 1e-4)``
(qio/grad/jacobi.py:139,149) and evaluates ``np.cos / np.sin`` of each angle (jacobi.py:36).
``t_opt`` can only be one of the 314 coarse points (the ``t_opt = 0 -> 0.01`` fallback of jacobi.py:146-147
equals coarse point 0), so 314 fine grids cover every case.  Building the angles and their cos/sin with
the very same numpy calls makes them bit-identical to the reference; the device only looks them up.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from ._lib import ThetaTables
from . import device as dev

COARSE_STEP = 0.01
FINE_STEP = 0.0001


def host_tables():
    coarse = np.arange(COARSE_STEP, np.pi, COARSE_STEP)
    fines = [np.arange(t - COARSE_STEP, t + COARSE_STEP, FINE_STEP) for t in coarse]
    stride = max(len(f) for f in fines)
    fine_len = np.array([len(f) for f in fines], dtype=np.int32)
    fine_theta = np.zeros((len(coarse), stride))
    for r, f in enumerate(fines):
        fine_theta[r, :len(f)] = f
    coarse_cs = np.stack([np.cos(coarse), np.sin(coarse)], axis=1)
    fine_cs = np.stack([np.cos(fine_theta), np.sin(fine_theta)], axis=2)
    return dict(coarse_theta=coarse, coarse_cs=np.ascontiguousarray(coarse_cs), fine_len=fine_len,
                fine_theta=fine_theta, fine_cs=np.ascontiguousarray(fine_cs), fine_stride=stride)


class DeviceTables:
    """Device copies + the C struct that points at them."""

    def __init__(self):
        h = host_tables()
        self.host = h
        self.coarse_theta = dev.to_device(h['coarse_theta'])
        self.coarse_cs = dev.to_device(h['coarse_cs'])
        self.fine_len = dev.to_device(h['fine_len'], dev.I32)
        self.fine_theta = dev.to_device(h['fine_theta'])
        self.fine_cs = dev.to_device(h['fine_cs'])
        self.struct = ThetaTables(len(h['coarse_theta']), int(h['fine_stride']),
                                  self.coarse_theta.data_ptr(), self.coarse_cs.data_ptr(), self.fine_len.data_ptr(),
                                  self.fine_theta.data_ptr(), self.fine_cs.data_ptr())


_CACHE: dict[int, DeviceTables] = {}


def device_tables() -> DeviceTables:
    d = dev.require_cuda()
    key = d.index if d.index is not None else torch.cuda.current_device()
    if key not in _CACHE:
        _CACHE[key] = DeviceTables()
    return _CACHE[key]
