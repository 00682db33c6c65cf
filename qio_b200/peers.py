"""Peer-mapped mailboxes for the multi-GPU Jacobi sweep (one process per GPU, one node).

Each rank cudaMalloc's a small mailbox (2 parities x nranks slots of 256 doubles + a tag), exports its CUDA IPC
handle, and maps every other rank's mailbox.  The persistent sweep kernel writes its partial super block straight
into the peers' mailboxes over NVLink and spins on the tags -- no NCCL call sits on the per-pair path.
``torch.distributed`` is used only to exchange the 64-byte handles (plumbing).
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import device as dev
from ._lib import Peers, check, lib


class PeerGroup:
    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerGroup needs torch.distributed to be initialised")
        dev.require_cuda()
        self.group = group
        self.rank = dist.get_rank(group)
        self.nranks = dist.get_world_size(group)
        nbytes = max(int(lib.qio_b200_mailbox_bytes(self.nranks)), int(lib.qio_b200_ll_stress_bytes(self.nranks)))
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        check(lib.qio_b200_ipc_alloc(nbytes, ctypes.byref(own), handle), 'ipc_alloc')
        self._own = own
        self._nbytes = nbytes
        self.wraps = 0
        handles = [None] * self.nranks
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._opened = []
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(own.value)
            else:
                p = ctypes.c_void_p()
                check(lib.qio_b200_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), 'ipc_open')
                self._opened.append(p)
                ptrs.append(p.value)
        self.ptr_table = torch.tensor(ptrs, dtype=torch.int64, device='cuda')
        self.struct = Peers(self.rank, self.nranks, 0, 0, self.ptr_table.data_ptr())
        dist.barrier(group=group)

    EPOCHS = 4095           # 12 tag bits (csrc/ll.cuh); 0 is never used

    def next_epoch(self):
        """Called once per sweep_cycle launch (identically on every rank).  The epoch distinguishes the launches in the
        mailbox tags; before a value is reused every rank zeroes its own mailbox, fenced by two barriers."""
        e = self.struct.epoch + 1
        if e > self.EPOCHS:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)          # nobody is still writing into anybody's mailbox
            check(lib.qio_b200_ipc_clear(self._own, self._nbytes, None), 'ipc_clear')
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            e = 1
            self.wraps += 1
        self.struct.epoch = e

    def stress(self, iterations=4000, lanes=256):
        """Self-test of the mailbox signalling: iterations x lanes tagged 16-byte stores from every rank into every
        rank's mailbox, every received payload verified.  Collective.  Returns (mismatches, timeouts, words_checked)."""
        self.next_epoch()
        res = torch.zeros(3, dtype=torch.int64, device='cuda')
        check(lib.qio_b200_ll_stress(ctypes.byref(self.struct), int(iterations), int(lanes), res.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream), 'll_stress')
        bad, lost, seen = (int(x) for x in res.cpu())
        return bad, lost, seen

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            lib.qio_b200_ipc_close(p)
        self._opened = []
        if self._own is not None:
            lib.qio_b200_ipc_free(self._own)
            self._own = None
