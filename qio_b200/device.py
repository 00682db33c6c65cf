"""Device plumbing: torch CUDA tensors as device-memory handles, the current torch stream, and thin
Python wrappers (one per C-ABI entry point) that pass raw pointers to libqio_b200.so.

PyTorch is used for allocation, streams and host<->device copies only -- every computation below
is a call into the hand-written CUDA library.  No function here has a CPU path.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import BLOCK_DOUBLES, LS_DTYPE, QioCudaError, check, lib

F64 = torch.float64
I32 = torch.int32


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise QioCudaError("qio_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device('cuda', torch.cuda.current_device())


def is_device(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def to_device(x, dtype=F64) -> torch.Tensor:
    """numpy / list / torch (any device) -> contiguous CUDA tensor of ``dtype`` (no copy if already so)."""
    dev = require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(x), dtype={F64: np.float64, I32: np.int32}[dtype])
    return torch.from_numpy(arr).to(dev, non_blocking=False)


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


def like_input(t: torch.Tensor, template):
    """Return ``t`` in the container type of ``template`` (numpy in -> numpy out, device in -> device out)."""
    return t if is_device(template) else to_host(t)


def _ptr(t) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def inactive_mask(inactive_indices, n: int) -> torch.Tensor:
    """int32 mask of length n from the reference's list of inactive orbitals."""
    m = np.zeros(n, dtype=np.int32)
    idx = np.asarray(list(inactive_indices), dtype=np.int64)
    if idx.size:
        if idx.min() < 0 or idx.max() >= n:
            raise IndexError("inactive index out of range")
        m[idx] = 1
    return to_device(m, I32)


def empty(*shape, dtype=F64) -> torch.Tensor:
    return torch.empty(*shape, dtype=dtype, device=require_cuda())


# ---- one wrapper per entry point -----------------------------------------------------------------
def prep_rdm12(dm1: torch.Tensor | None, dm2: torch.Tensor, n: int, na: int | None = None,
               gamma: torch.Tensor | None = None, Gamma: torch.Tensor | None = None):
    na = n if na is None else na
    if dm1 is not None and gamma is None:
        gamma = empty(2 * n, 2 * n)
    if Gamma is None:
        Gamma = empty(na, n, n, n)
    check(lib.qio_b200_prep_rdm12(_ptr(dm1), _ptr(dm2), _ptr(gamma if dm1 is not None else None), _ptr(Gamma),
                                   n, na, _stream()), 'prep_rdm12')
    return gamma, Gamma


def orbital_diagonals(gamma, Gamma, n, inds: torch.Tensor, a0=0, na=None) -> torch.Tensor:
    na = n if na is None else na
    m = int(inds.numel())
    diag = empty(3, m)
    check(lib.qio_b200_orbital_diagonals(_ptr(gamma), _ptr(Gamma), n, a0, na, _ptr(inds), m, _ptr(diag), _stream()),
          'orbital_diagonals')
    return diag


def orbital_entropies(diag: torch.Tensor, want_spec=True, want_unmasked=False):
    m = int(diag.shape[1])
    spec = empty(4, m) if want_spec else None
    s_masked = empty(m)
    s_un = empty(m) if want_unmasked else None
    summary = empty(4)
    check(lib.qio_b200_orbital_entropies(_ptr(diag), m, _ptr(spec), _ptr(s_masked), _ptr(s_un), _ptr(summary),
                                          _stream()), 'orbital_entropies')
    return spec, s_masked, s_un, summary


def shannon(p: torch.Tensor) -> torch.Tensor:
    summary = empty(4)
    check(lib.qio_b200_shannon(_ptr(p), int(p.numel()), _ptr(summary), _stream()), 'shannon')
    return summary


def pair_blocks(gamma, Gamma, n, pairs: torch.Tensor, a0=0, na=None) -> torch.Tensor:
    na = n if na is None else na
    P = int(pairs.shape[0])
    blocks = empty(P, BLOCK_DOUBLES)
    check(lib.qio_b200_pair_blocks(_ptr(gamma), _ptr(Gamma), n, a0, na, _ptr(pairs), P, _ptr(blocks), _stream()),
          'pair_blocks')
    return blocks


def fqi_grad_hess(blocks, n, mask):
    dX, ddX = empty(n, n), empty(n, n)
    check(lib.qio_b200_fqi_grad_hess(_ptr(blocks), n, _ptr(mask), _ptr(dX), _ptr(ddX), _stream()), 'fqi_grad_hess')
    return dX, ddX


def pair_grad_hess(blocks, pairs, mask):
    P = int(pairs.shape[0])
    g1, g2 = empty(P), empty(P)
    check(lib.qio_b200_pair_grad_hess(_ptr(blocks), _ptr(pairs), P, _ptr(mask), _ptr(g1), _ptr(g2), _stream()),
          'pair_grad_hess')
    return g1, g2


def fqi_grad_hess_fused(gamma, Gamma, n, mask):
    dX, ddX = empty(n, n), empty(n, n)
    check(lib.qio_b200_fqi_grad_hess_fused(_ptr(gamma), _ptr(Gamma), n, _ptr(mask), _ptr(dX), _ptr(ddX), _stream()),
          'fqi_grad_hess_fused')
    return dX, ddX


def pair_cost(blocks, cand_block, cand_pair, cand_cs, mask):
    C = int(cand_block.numel())
    cost = empty(C)
    flags = empty(C, dtype=I32)
    check(lib.qio_b200_pair_cost(_ptr(blocks), _ptr(cand_block), _ptr(cand_pair), _ptr(cand_cs), C, _ptr(mask),
                                  _ptr(cost), _ptr(flags), _stream()), 'pair_cost')
    return cost, flags


def pair_linesearch(blocks, pairs, mask, tables, strict_order=False) -> torch.Tensor:
    P = int(pairs.shape[0])
    results = torch.empty(P * LS_DTYPE.itemsize, dtype=torch.uint8, device=require_cuda())
    check(lib.qio_b200_pair_linesearch(_ptr(blocks), _ptr(pairs), P, _ptr(mask), ctypes.byref(tables.struct),
                                        _ptr(results), 1 if strict_order else 0, _stream()), 'pair_linesearch')
    return results


def results_to_numpy(results: torch.Tensor) -> np.ndarray:
    return results.cpu().numpy().view(LS_DTYPE)


def givens_apply(gamma, Gamma, n, i, j, c, s, zero_offspin=False):
    check(lib.qio_b200_givens_apply(_ptr(gamma), _ptr(Gamma), n, int(i), int(j), float(c), float(s),
                                     1 if zero_offspin else 0, _stream()), 'givens_apply')


def rotate_rdm1(U, gamma, n) -> torch.Tensor:
    out = torch.empty_like(gamma)
    check(lib.qio_b200_rotate_rdm1(_ptr(U), _ptr(gamma), _ptr(out), n, _stream()), 'rotate_rdm1')
    return out


def rotate_rdm2_(U, Gamma, n, work=None) -> torch.Tensor:
    """In place (needs an n^4 work buffer; allocated when not supplied)."""
    if work is None:
        work = torch.empty_like(Gamma)
    check(lib.qio_b200_rotate_rdm2(_ptr(U), _ptr(Gamma), _ptr(work), n, _stream()), 'rotate_rdm2')
    return Gamma


def rotate_axis(U, src, dst, n, rows, ld_in):
    check(lib.qio_b200_rotate_axis(_ptr(U), _ptr(src), _ptr(dst), n, int(rows), int(ld_in), _stream()), 'rotate_axis')
    return dst


def rotate_axis_batched(U, src, dst, n, rows, ld_in, batch, bs_in, bs_out):
    """``batch`` cyclic single-axis passes (see rotate_axis) in one launch."""
    check(lib.qio_b200_rotate_axis_batched(_ptr(U), _ptr(src), _ptr(dst), n, int(rows), int(ld_in), int(batch), int(bs_in),
                                            int(bs_out), _stream()), 'rotate_axis_batched')
    return dst


def jacobi_cycle(gamma, Gamma, U, n, pairs, mask, tables):
    P = int(pairs.shape[0])
    results = torch.empty(max(P, 1) * LS_DTYPE.itemsize, dtype=torch.uint8, device=require_cuda())
    summary = empty(4)
    ws_bytes = int(lib.qio_b200_jacobi_workspace_bytes(n, P))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=require_cuda())
    check(lib.qio_b200_jacobi_cycle(_ptr(gamma), _ptr(Gamma), _ptr(U), n, _ptr(pairs), P, _ptr(mask),
                                     ctypes.byref(tables.struct), _ptr(results), _ptr(summary), _ptr(ws), _stream()),
          'jacobi_cycle')
    return results[:P * LS_DTYPE.itemsize], summary


def contract_axis(U, src, dst, n, rows, last_axis):
    check(lib.qio_b200_contract_axis(_ptr(U), _ptr(src), _ptr(dst), n, int(rows), 1 if last_axis else 0, _stream()),
          'contract_axis')
    return dst


def contract_axis_cyclic(U, src, dst, n_out, kdim, rows):
    """dst[m, i] = sum_{k<kdim} U[i, k] src[k, m]: src (kdim, rows) -> dst (rows, n_out), U (n_out, kdim) row-major."""
    check(lib.qio_b200_contract_axis_cyclic(_ptr(U), int(kdim), _ptr(src), _ptr(dst), int(n_out), int(kdim), int(rows), _stream()),
          'contract_axis_cyclic')
    return dst


def contract_axis_partial(U, ldu, src, dst, n, kdim, rows):
    """dst[i, m] = sum_{k<kdim} U[i*ldu + k] src[k, m] (U may be a column block of a larger matrix)."""
    check(lib.qio_b200_contract_axis_partial(_ptr(U), int(ldu), _ptr(src), _ptr(dst), n, int(kdim), int(rows), _stream()),
          'contract_axis_partial')
    return dst


def sweep_reset(W, n):
    check(lib.qio_b200_sweep_reset(_ptr(W), n, _stream()), 'sweep_reset')


def sweep_cycle(gamma, S, U, W, n, pairs, mask, tables, a0=0, na=None, peers=None, engine=0, timers=False,
                blocks_out=None, timeout_ms=0, startup_timeout_ms=0, trace_out=None, trace_first_pair=0):
    """One cycle on the deferred-axis state (S, W); returns (results bytes, summary).
    ``peers`` is a qio_b200.peers.PeerGroup (multi-GPU) or None.  ``engine``: 0 = automatic, 2 = barrier kernel,
    3 = pipelined kernel; ``timers``: the build that carries phase timers; ``blocks_out``: (P, 24) device tensor that
    receives the pair block every line search ran on (see audit_decisions).  All of it travels in the call's options
    struct -- the library keeps no state and reads no environment variables."""
    na = n if na is None else na
    P = int(pairs.shape[0])
    results = torch.empty(max(P, 1) * LS_DTYPE.itemsize, dtype=torch.uint8, device=require_cuda())
    summary = empty(32)
    ws = torch.empty(int(lib.qio_b200_sweep_workspace_bytes(n)), dtype=torch.uint8, device=require_cuda())
    if peers is not None:
        peers.next_epoch()
    opts = _lib.SweepOptions(int(engine), 1 if timers else 0, int(timeout_ms), int(startup_timeout_ms),
                             0 if blocks_out is None else blocks_out.data_ptr(),
                             0 if trace_out is None else trace_out.data_ptr(), int(trace_first_pair), 0)
    check(lib.qio_b200_sweep_cycle(_ptr(gamma), _ptr(S), _ptr(U), _ptr(W), n, a0, na, _ptr(pairs), P, _ptr(mask),
                                    ctypes.byref(tables.struct), _ptr(results), _ptr(summary), _ptr(ws),
                                    ctypes.byref(peers.struct) if peers is not None else None, ctypes.byref(opts),
                                    _stream()), 'sweep_cycle')
    return results[:P * LS_DTYPE.itemsize], summary


def audit_decisions(records: np.ndarray, blocks, pairs, mask, tables, near_tie=1e-12):
    """Re-derives every line-search decision of a cycle from the pair blocks the sweep kernel itself searched
    (``blocks_out`` of sweep_cycle), in the reference's own floating-point operation order (strict_order: 16-term sums,
    library log -- qio/grad/jacobi.py:15-60,117-158), as one batched launch, and compares it with the decisions the kernel
    took with its polynomial cost.  Returns a dict: pairs, mismatches (pair indices whose coarse / fine index differ),
    near_ties (pairs whose smaller margin is below ``near_tie``), unexplained (mismatches that are NOT near-ties: those
    would be a real defect), min_fine_margin, min_coarse_margin."""
    P = int(pairs.shape[0])
    if P == 0:
        return dict(pairs=0, mismatches=[], near_ties=[], unexplained=[], min_fine_margin=float('inf'),
                    min_coarse_margin=float('inf'))
    strict = results_to_numpy(pair_linesearch(blocks, pairs, mask, tables, strict_order=True))
    diff = (strict['coarse_index'] != records['coarse_index']) | (strict['fine_index'] != records['fine_index'])
    # a pair neither of whose orbitals is inactive has an identically zero cost: no decision to audit
    fm = np.where(np.isfinite(records['fine_margin']), records['fine_margin'], np.inf)
    cm = np.where(np.isfinite(records['coarse_margin']), records['coarse_margin'], np.inf)
    sfm = np.where(np.isfinite(strict['fine_margin']), strict['fine_margin'], np.inf)
    scm = np.where(np.isfinite(strict['coarse_margin']), strict['coarse_margin'], np.inf)
    margin = np.minimum(np.minimum(fm, cm), np.minimum(sfm, scm))
    tie = margin < near_tie
    # the pipelined sweep kernel leaves its own margins at +inf (off its critical path): the strict pass supplies them
    return dict(pairs=P, mismatches=np.nonzero(diff)[0].tolist(), near_ties=np.nonzero(tie)[0].tolist(),
                unexplained=np.nonzero(diff & ~tie)[0].tolist(), min_fine_margin=float(np.minimum(fm, sfm).min()),
                min_coarse_margin=float(np.minimum(cm, scm).min()))


def sweep_engine(n, na=None):
    return int(lib.qio_b200_sweep_engine(n, n if na is None else na))


def check_sweep_summary(summary_host):
    """Raises if the pipelined sweep kernel reported a timed-out wait (summary[3]).  The cycle is then incomplete and
    S, W, gamma, U are partly rotated: callers mark their state as poisoned and refuse further cycles / flushes."""
    if float(summary_host[3]) != 0.0:
        raise QioCudaError('sweep_cycle: a wait inside the pipelined kernel timed out; the cycle is incomplete')


def deferred_diagonals(gamma, S, W, n, inds, a0=0, na=None):
    na = n if na is None else na
    m = int(inds.numel())
    diag = empty(3, m)
    check(lib.qio_b200_deferred_diagonals(_ptr(gamma), _ptr(S), _ptr(W), n, a0, na, _ptr(inds), m, _ptr(diag),
                                           _stream()), 'deferred_diagonals')
    return diag


def zero_offspin_if_rotated(gamma, n, summary):
    check(lib.qio_b200_zero_offspin_if_rotated(_ptr(gamma), n, _ptr(summary), _stream()), 'zero_offspin_if_rotated')


def sweep_flush(S, W, n, work=None):
    """Gamma = W applied to axes 0 and 3 of S (two DMMA GEMM passes), result in S; W reset to 1."""
    if work is None:
        work = torch.empty_like(S)
    n3 = n ** 3
    contract_axis(W, S, work, n, n3, last_axis=False)
    contract_axis(W, work, S, n, n3, last_axis=True)
    sweep_reset(W, n)
    return S


def pair_entropy(blocks, supp=None):
    """(s_pair (P), eig (P, 16), flags (P)) of the two-orbital RDMs built from the pair blocks (extension a-9)."""
    P = int(blocks.shape[0])
    s_pair, eig, flags = empty(P), empty(P, 16), empty(P, dtype=I32)
    check(lib.qio_b200_pair_entropy(_ptr(blocks), _ptr(supp), P, _ptr(s_pair), _ptr(eig), _ptr(flags), _stream()),
          'pair_entropy')
    return s_pair, eig, flags


def mutual_information_matrix(s1, s_pair, pairs, n):
    """(n, n) device matrix I_ij = S_i + S_j - S_ij for ``pairs`` (P, 2); zero elsewhere (extension a-9)."""
    I = empty(n, n)
    check(lib.qio_b200_mutual_information(_ptr(s1), _ptr(s_pair), _ptr(pairs), int(pairs.shape[0]), n, _ptr(I), _stream()),
          'mutual_information')
    return I
