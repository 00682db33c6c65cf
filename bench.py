#!/usr/bin/env python
"""Benchmark of the QIO orbital-entropy hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--orbitals 100] [--impl reference] [--no-north-star]

One "step" is one complete pass of the hot path over one synthetic solver output:

    prep_rdm12(dm1, dm2)                        a-1   (restores the pristine RDMs every step)
    random-start orbital rotation (2- and 4-index, FP64 DMMA)   a-7 start / a-10 rotation
    all-pairs pass: pair blocks, gradient + Hessian, angle line search for all n(n-1)/2 pairs, pair entropies and the
        mutual-information matrix                                                           a-8, a-5, a-9
    one full Jacobi cycle: 4950 (n=100) sequential line searches + Givens updates of the RDMs, followed by the
        strict-order audit of every decision                                                a-7, a-6
    get_cost_fqi of the result, flush of the deferred state      a-3

metric  = orbital-pair entropy evaluations (jacobi_cost calls of the reference) per second over the step;
ms_per_step is the wall time of the whole step (all-pairs pass + Jacobi sweep).
`value`: inputs (dm1, dm2, start rotation, pair order) resident in HBM.   `e2e`: the same step through the
public Python API from pinned HOST buffers, H2D of dm1/dm2 and D2H of the results inside the timed region.
`north_star`: the same two measurements on BASELINE.json configs[4] (n = 200), at whatever GPU count the run has.
"""
from __future__ import annotations

import os
import sys

if '--impl' in sys.argv and 'reference' in sys.argv:
    # The CPU arm uses every host core for its BLAS part whatever the launcher exported (torchrun sets
    # OMP_NUM_THREADS=1 for its children): fix the thread counts BEFORE numpy loads its BLAS.
    _cores = str(os.cpu_count() or 1)
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ[_v] = _cores

import argparse
import importlib.util
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'orbital_pair_entropy_evals_per_s'
UNIT = 'evals/s'
NORTH_STAR_N = 200


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    # '--orbitals' is the spelling to use under torchrun: its own parser rejects '--n' as an ambiguous prefix
    ap.add_argument('--n', '--orbitals', dest='n', type=int, default=int(os.environ.get('QIO_BENCH_N', '100')),
                    help='number of orbitals of the headline leg (100 = configs[3])')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-north-star', action='store_true', help='skip the second leg on configs[4] (n = 200)')
    ap.add_argument('--no-parity', action='store_true', help='N > 1: skip the sharded-vs-single-GPU self-check')
    return ap.parse_args()


def load_synth():
    """qio_b200/synth.py as a stand-alone module: plain numpy, and importing it this way does not load the CUDA library
    (the reference arm must not map the product's .so)."""
    spec = importlib.util.spec_from_file_location('qio_b200_synth_standalone', os.path.join(ROOT, 'qio_b200', 'synth.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# --------------------------------------------------------------------------------------------------
# workload description shared by both arms
# --------------------------------------------------------------------------------------------------
def workload_config(n, n_gpus):
    return {
        'workload': f'synthetic determinant-ensemble rdm1/rdm2, n={n} orbitals (K=4, nocc={max(1, n // 5)}, seed 20261019+n), '
                    f'inactive = all orbitals: prep_rdm12 + random-start rotation + all-pairs grad/Hess/line-search/mutual-information '
                    f'+ one full Jacobi cycle',
        'n_orbitals': n,
        'pairs': n * (n - 1) // 2,
        'numpy_seed': 0,
        'parallelism': f'{n_gpus} gpu' if n_gpus == 1 else f'rdm2 sharded by leading index over {n_gpus} gpus',
        'l2_policy': 'inputs larger than L2 (rdm2 = %.2f GB, rewritten by prep_rdm12 every step)' % (8 * n ** 4 / 1e9),
    }


def sweep_order(n):
    """The reference's RNG call order under np.random.seed(0): rand(n,n), rand(), shuffle (jacobi.py:197,214)."""
    from scipy.linalg import expm
    np.random.seed(0)
    X = (np.random.rand(n, n) / 2 - 1) / 1 / (1 + 49 * np.random.rand())
    X = X - X.T
    U0 = expm(X)
    order = np.arange(0, n)
    np.random.shuffle(order)
    return U0, order


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def wait_ready(self, timeout=6.0):
        """Block until the first sample has arrived: nvidia-smi's start-up (NVML attaches to every GPU of the box) can stall
        launches on all of them for tens of milliseconds and must be over before the timed region begins."""
        t0 = time.time()
        while self.proc is not None and not self.rows and self.proc.poll() is None and time.time() - t0 < timeout:
            time.sleep(0.02)
        time.sleep(0.1)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(smax)), 'power_w_max': float(max(power)),
                'samples': len(sm), 'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------------------------------
_CPU_SETUP = {}


def cpu_functions():
    """(kind, line search, transform, prepared-RDM rotation helpers).  'reference' = the unmodified reference package copied
    to oracle/_ref by oracle/make_ref.py; 'port' = the oracle restatement (only when that copy does not exist)."""
    from oracle import make_ref
    from oracle import qio_oracle as O
    ref = make_ref.load_reference()
    if ref is not None:
        return 'reference', ref['jacobi'].jacobi_direct, ref['jacobi'].jacobi_transform, O
    return 'port', O.jacobi_direct_scalar, O.jacobi_transform_dense, O


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get('num_threads', 1)) for p in threadpool_info() if p.get('user_api') == 'blas'] or [1])
    except Exception:       # noqa: BLE001 (diagnostic only)
        return None


def cpu_reference_setup(n):
    """Rotated synthetic RDMs and the sweep pair order on the host (built once per process; plain numpy)."""
    if n not in _CPU_SETUP:
        from oracle import qio_oracle as O
        synth = load_synth()
        t0 = time.time()
        gamma0, Gamma0 = synth.prepared_rdms(n)
        U0, order = sweep_order(n)
        g = O.rotate_rdm1(U0, gamma0)
        G = O.rotate_rdm2(U0, Gamma0)
        inact = list(range(n))
        _CPU_SETUP[n] = dict(g=g, G=G, inact=inact, pairs=O.sweep_pairs(order, inact), setup_s=time.time() - t0)
    return _CPU_SETUP[n]


def cpu_reference_sample(n, n_ls_pairs=48, n_sweep_pairs=4, offset=0):
    """Times the reference's CPU path (qio/grad/jacobi.py: the pure-Python jacobi_cost loop of jacobi_direct and the dense
    einsum of jacobi_transform) on a bounded sample and extrapolates the evals/s of the full step:
    all-pairs pass = P line searches;  sweep = P x (line search + accepted transform)."""
    kind, line_search, transform, _ = cpu_functions()
    st = cpu_reference_setup(n)
    g, G, inact, pairs = st['g'], st['G'], st['inact'], st['pairs']
    P = len(pairs)
    sample = [pairs[(offset + k) % P] for k in range(n_ls_pairs)]
    # line searches (1 core: pure Python in the reference)
    t0 = time.time()
    thetas = [line_search(i, j, g, G, inact) for (i, j) in sample]
    t_ls = (time.time() - t0) / n_ls_pairs
    accepted = sum(t is not None for t in thetas)
    # dense transforms (BLAS threads)
    t0 = time.time()
    done = 0
    for (i, j), t in zip(sample[:n_sweep_pairs], thetas[:n_sweep_pairs]):
        transform(g, G, i, j, 0.1 if t is None else t)
        done += 1
    t_tr = (time.time() - t0) / max(done, 1)
    acc_rate = accepted / n_ls_pairs
    evals_per_pair = 515.5
    step_s = P * t_ls + P * (t_ls + acc_rate * t_tr)
    value = 2 * P * evals_per_pair / step_s
    nblas = blas_threads()
    what = ("the reference's own qio.grad.jacobi (unmodified copy under oracle/_ref)" if kind == 'reference'
            else 'oracle port of the reference CPU path (oracle/_ref missing)')
    return {
        'value': value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': kind, 'blas_threads': nblas,
        'sample': (f'{what} at n={n}: {n_ls_pairs} jacobi_direct line searches '
                   f'({t_ls * 1e3:.1f} ms each, 1 core, pure Python) + {done} jacobi_transform einsums ({t_tr:.2f} s each, '
                   f'BLAS on {nblas} of {os.cpu_count()} host threads, OMP_NUM_THREADS={os.environ.get("OMP_NUM_THREADS", "unset")}); '
                   f'full step EXTRAPOLATED as P*t_ls + P*(t_ls + {acc_rate:.2f}*t_tr) = {step_s:.0f} s for P={P} pairs'),
        'line_search_ms': t_ls * 1e3, 'transform_s': t_tr, 'extrapolated_step_s': step_s, 'setup_s': st['setup_s'],
    }


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, args.steps)
    vals = []
    last = None
    t_start = time.time()
    for s in range(args.warmup + steps):
        # bounded sample per step so that the whole run stays within minutes
        last = cpu_reference_sample(args.n, n_ls_pairs=48, n_sweep_pairs=2, offset=48 * s)
        if s >= args.warmup:
            vals.append(last['value'])
        if time.time() - t_start > 240 and vals:
            break
    value = float(np.mean(vals))
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(vals),
        'warmup': args.warmup, 'ms_per_step': last['extrapolated_step_s'] * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args.n, args.gpus),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': last['cores'], 'kind': last['kind'], 'sample': last['sample'],
                         'blas_threads': last['blas_threads']},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'extrapolated': True,
        'loaded_product_library': any('libqio_b200' in line for line in open('/proc/self/maps')),
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def moved_bytes(n, na, pairs_np, accepted, look=3):
    """Bytes the deferred layout really moves in the flat + tile phases of a cycle: per accepted rotation w the rows
    (i|j, x) and (x, i|j) of every ACTIVE orbital x (one that has appeared in a super block <= w + look) over the na local
    slabs, read + written: 2 axes x 2 rows x n_active x n x na x 8 B x 2."""
    seen = np.zeros(n, dtype=bool)
    active_at = np.empty(len(pairs_np), dtype=np.int64)
    first = {}
    for k, (i, j) in enumerate(pairs_np):
        for x in (int(i), int(j)):
            first.setdefault(x, k)
    first_arr = np.array([first.get(x, 1 << 60) for x in range(n)])
    for w in range(len(pairs_np)):
        active_at[w] = int((first_arr <= w + look).sum())
    acc = np.asarray(accepted, dtype=bool)
    return float((2 * 2 * active_at[acc] * n * na * 8 * 2).sum())


def count_launches(engine, world=1, audit=True):
    """Kernels of THIS library launched by one device step (torch copies / NCCL calls are not counted)."""
    c = 0
    c += 2                  # prep_rdm12: 1-RDM interleave + 2-RDM permute/combine
    c += 1 + (4 if world == 1 else 3 + 1)   # rotate_rdm1; rotate_rdm2: 4 passes, or 3 batched passes + the column-block GEMM
    c += 1 + 1 + 1 + 1      # pair_blocks, grad/Hess, line search of all pairs, pair entropies
    c += 1 + 1 + 1          # mutual-information matrix: diagonals, one-orbital entropies, assembly
    c += 1                  # sweep_reset
    c += 1 + (1 if engine == 3 else 0)      # sweep kernel (+ its per-pair table kernel)
    c += 1 + 1 + 1          # deferred diagonals, entropies (cost), off-spin zeroing
    c += 1 if audit else 0  # strict-order re-evaluation of every decision
    c += 2 + 1              # flush: last-axis pass + leading-axis pass, W reset
    return c


def run_b200_arm(args):
    # NCCL prints its version banner (and, at INFO, its whole log) on STDOUT: default to WARN so that rank 0's JSON line is
    # the only thing there; a caller who exports NCCL_DEBUG keeps it.  The communicator size is reported by the bench itself
    # (stderr line per rank + "comm_nranks" in the JSON line), and the JSON line is printed last, after the group is destroyed.
    os.environ.setdefault('NCCL_DEBUG', 'WARN')
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        args.gpus = world
        return run_b200_multi(args, rank, world, local_rank)

    out = single_gpu_leg(args, args.n, args.steps, max(args.warmup, 3), local_rank, detail=True)
    if not args.no_north_star and args.n != NORTH_STAR_N:
        torch.cuda.empty_cache()
        ns = single_gpu_leg(args, NORTH_STAR_N, max(1, min(args.steps, 2)), 3, local_rank, detail=False)
        out['north_star'] = north_star_record(ns)
    print(json.dumps(out))


def north_star_record(leg):
    """BASELINE.json configs[4] as a sub-record of the JSON line."""
    keep = ('ms_per_step', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'e2e', 'roofline', 'clocks', 'rotations_accepted',
            'cost_after_cycle', 'decision_audit', 'phases', 'evals_per_step', 'peak_memory_gb')
    rec = {k: leg.get(k) for k in keep}
    rec['n'] = leg['config']['n_orbitals']
    rec['config'] = leg['config']
    rec['target'] = 'mutual-information matrix + one complete Jacobi sweep at n=200 in < 1 s on 8 B200 (BASELINE.json north_star)'
    rec['meets_target_ms'] = bool(leg['ms_per_step'] < 1000.0)
    return rec


def single_gpu_leg(args, n, steps, warmup, local_rank, detail):
    import torch

    from qio_b200 import device as dev
    from qio_b200 import mutual_info as MI
    from qio_b200 import synth, tables
    from qio_b200._lib import LS_DTYPE
    from qio_b200.grad import jacobi as JA
    from qio_b200.qio import prep_rdm12

    P = n * (n - 1) // 2
    inact = list(range(n))
    # ---- synthetic solver output on the host (pinned) and on the device ---------------------------
    dm1_h = torch.empty((n, n), dtype=torch.float64).pin_memory()
    dm2_h = torch.empty((n, n, n, n), dtype=torch.float64).pin_memory()
    w, Pk = synth.projectors(n)
    d1 = 2.0 * torch.from_numpy(Pk).cuda()
    wt = torch.from_numpy(w).cuda()
    dm1_d = torch.einsum('k,kab->ab', wt, d1).contiguous()
    dm2_d = torch.zeros((n, n, n, n), dtype=torch.float64, device='cuda')
    for k in range(len(w)):     # input synthesis only (plumbing, outside every timed region), slab-wise to bound the temporaries
        for a0 in range(0, n, 25):
            dl = d1[k][a0:a0 + 25]
            dm2_d[a0:a0 + 25] += wt[k] * (torch.einsum('pq,rs->pqrs', dl, d1[k]) - 0.5 * torch.einsum('ps,rq->pqrs', dl, d1[k]))
    dm1_h.copy_(dm1_d)
    dm2_h.copy_(dm2_d)
    torch.cuda.synchronize()

    U0, order = sweep_order(n)
    pairs_np = JA.sweep_pairs(order, inact)
    a, b = np.tril_indices(n, -1)
    allpairs_np = np.stack([a, b], 1).astype(np.int32)
    U0_d = dev.to_device(U0)
    pairs_d = dev.to_device(pairs_np, dev.I32)
    allpairs_d = dev.to_device(allpairs_np, dev.I32)
    mask = dev.inactive_mask(inact, n)
    tabs = tables.device_tables()
    fine_len = tabs.host['fine_len']

    Gamma = dev.empty(n, n, n, n)
    gamma = dev.empty(2 * n, 2 * n)
    work = dev.empty(n, n, n, n)
    W = dev.empty(n, n)
    blocks_sweep = dev.empty(len(pairs_np), dev.BLOCK_DOUBLES)
    inds_d = dev.to_device(np.arange(n, dtype=np.int32), dev.I32)
    state = {}
    engine = dev.sweep_engine(n)
    torch.cuda.reset_peak_memory_stats()

    sweep_events = []           # CUDA events around the sweep kernel of every device_step

    def device_step():
        """The whole hot path with every input already in HBM."""
        dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma)
        g = dev.rotate_rdm1(U0_d, gamma, n)
        dev.rotate_rdm2_(U0_d, Gamma, n, work)
        blocks = dev.pair_blocks(g, Gamma, n, allpairs_d)
        dX, ddX = dev.fqi_grad_hess(blocks, n, mask)
        ls_all = dev.pair_linesearch(blocks, allpairs_d, mask, tabs)
        s_pair, _, _ = dev.pair_entropy(blocks)          # extension a-9: pair entropies ...
        _, s1, _, _ = dev.orbital_entropies(dev.orbital_diagonals(g, Gamma, n, inds_d), want_spec=False)
        mi = dev.mutual_information_matrix(s1, s_pair, allpairs_d, n)      # ... and the mutual-information matrix
        Ud = U0_d.clone()
        dev.sweep_reset(W, n)
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
        res, summary = dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs, blocks_out=blocks_sweep)
        ev[1].record()
        sweep_events.append(ev)
        strict = dev.pair_linesearch(blocks_sweep, pairs_d, mask, tabs, strict_order=True)      # decision audit
        diag = dev.deferred_diagonals(g, Gamma, W, n, inds_d)
        _, _, _, esum = dev.orbital_entropies(diag, want_spec=False)
        dev.zero_offspin_if_rotated(g, n, summary)
        dev.sweep_flush(Gamma, W, n, work)
        state.update(ls_all=ls_all, res=res, strict=strict, summary=summary, esum=esum, U=Ud, dX=dX, g=g, mi=mi)

    def count_evals():
        ra = dev.results_to_numpy(state['ls_all'])
        rs = dev.results_to_numpy(state['res'])
        rows_a = np.where(ra['coarse_index'] >= 0, ra['coarse_index'], 0)
        rows_s = np.where(rs['coarse_index'] >= 0, rs['coarse_index'], 0)
        ev = int((1 + 314 + fine_len[rows_a]).sum() + (1 + 314 + fine_len[rows_s]).sum())
        return ev, int(rs['accepted'].sum()), rs

    def api_step():
        """The same step through the public API from pinned host buffers (e2e)."""
        g0, G0 = prep_rdm12(dm1_h.numpy(), dm2_h, on_device=True)
        rec_all = JA.linesearch_pairs(g0, G0, allpairs_np, inact)
        mi = MI.mutual_information(g0, G0)
        np.random.seed(0)
        rots, U, g1, G1 = JA.minimize_orb_corr_jacobi(g0, G0, inact, 1)      # includes the decision audit (JA.AUDIT)
        mi_h = dev.to_host(mi)
        torch.cuda.synchronize()
        return rec_all, rots, U, mi_h

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(warmup):
        device_step()
    torch.cuda.synchronize()
    dev.check_sweep_summary(dev.to_host(state['summary']))
    evals_per_step, n_accepted, rs = count_evals()
    cost_after = float(dev.to_host(state['esum'])[0])
    audit = dev.audit_decisions(rs, blocks_sweep, pairs_d, mask, tabs)
    decision_audit = {'pairs': audit['pairs'], 'mismatches': len(audit['mismatches']), 'near_ties_below_1e-12': len(audit['near_ties']),
                      'unexplained': len(audit['unexplained']), 'min_fine_margin': audit['min_fine_margin'],
                      'min_coarse_margin': audit['min_coarse_margin'],
                      'how': 'every line-search decision of the cycle re-derived in the reference operation order (strict_order) from the '
                             '24-number block the sweep kernel searched; same coarse and fine index required'}
    if audit['unexplained']:
        raise RuntimeError(f'decision audit failed: {audit["unexplained"][:5]}')

    # ---- timed: device-resident -----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_ready()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    del sweep_events[:]
    e0.record()
    for _ in range(steps):
        device_step()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    sweep_in_step = [a.elapsed_time(b) for a, b in sweep_events]      # the dominant kernel inside the timed region
    clocks = sampler.stop()
    ms_per_step = ms_total / steps
    value = evals_per_step / (ms_per_step * 1e-3)
    launches_per_step = count_launches(engine)
    peak_mem = torch.cuda.max_memory_allocated() / 1e9

    # ---- phase timings (explain the step) and the rooflines -------------------------------------------------
    def timed(fn, reps=1):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a_.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / reps

    peak, peak_src = hbm_peak()
    phases = {}
    phases['prep_rdm12_ms'] = timed(lambda: dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma), 3)
    phases['rotate_rdm2_ms'] = timed(lambda: dev.rotate_rdm2_(U0_d, Gamma, n, work), 3)
    g = dev.rotate_rdm1(U0_d, gamma, n)
    blocks = dev.pair_blocks(g, Gamma, n, allpairs_d)
    phases['pair_blocks_ms'] = timed(lambda: dev.pair_blocks(g, Gamma, n, allpairs_d), 5)
    phases['grad_hess_ms'] = timed(lambda: dev.fqi_grad_hess(blocks, n, mask), 5)
    phases['linesearch_all_pairs_ms'] = timed(lambda: dev.pair_linesearch(blocks, allpairs_d, mask, tabs), 5)
    phases['pair_entropy_ms'] = timed(lambda: dev.pair_entropy(blocks), 5)
    s_pair = dev.pair_entropy(blocks)[0]
    phases['mutual_information_matrix_ms'] = timed(lambda: dev.mutual_information_matrix(
        dev.orbital_entropies(dev.orbital_diagonals(g, Gamma, n, inds_d), want_spec=False)[1], s_pair, allpairs_d, n), 5)
    phases['decision_audit_ms'] = timed(lambda: dev.pair_linesearch(blocks_sweep, pairs_d, mask, tabs, strict_order=True), 3)
    m1_evals = int((1 + 314 + fine_len[np.where(dev.results_to_numpy(state['ls_all'])['coarse_index'] >= 0,
                                                dev.results_to_numpy(state['ls_all'])['coarse_index'], 0)]).sum())
    phases['all_pairs_evals_per_s'] = m1_evals / (phases['linesearch_all_pairs_ms'] * 1e-3)

    # dominant kernel = the persistent sweep kernel, timed alone on the step's own first cycle (same start state)
    def fresh_state():
        dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma)
        g_ = dev.rotate_rdm1(U0_d, gamma, n)
        dev.rotate_rdm2_(U0_d, Gamma, n, work)
        dev.sweep_reset(W, n)
        return g_, U0_d.clone()

    g, Ud = fresh_state()
    phases['jacobi_cycle_kernel_alone_ms'] = timed(lambda: dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs), 1)
    sweep_ms = float(np.mean(sweep_in_step))        # average launch duration over the timed steps (CUDA events, same stream)
    phases['jacobi_cycle_kernel_ms'] = sweep_ms
    phases['jacobi_cycle_kernel_ms_each'] = sweep_in_step
    phases['sweep_flush_ms'] = timed(lambda: dev.sweep_flush(Gamma, W, n, work), 1)
    if detail:
        g, Ud = fresh_state()
        trace = dev.empty(32, 16)
        tsum = dev.to_host(dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs, timers=True, trace_out=trace,
                                           trace_first_pair=len(pairs_np) // 2)[1])   # build with phase timers (~5 % slower)
        phases['sweep_phase_us_per_pair'] = sweep_phase_table(tsum, len(pairs_np))
        phases['sweep_chain_us'] = chain_table(dev.to_host(trace))
        lsn = ['coarse', 'reduce', 'fine', 'scan', 'margin', 'tail']
        phases['linesearch_cycles_per_pair_cta0'] = {k: float(v) / len(pairs_np) for k, v in zip(lsn, tsum[10:16])}
        # one Newton-Raphson ("nr", the path of the reference's examples) iteration through the public API, device-resident
        # RDMs: grad/Hess kernels, D2H of two n x n matrices, host level shift + scipy expm, H2D of U, 4-index rotation, cost
        from qio_b200.grad import gradient as GR
        dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma)
        GR.minimize_orb_corr_GD(gamma, Gamma, inact, np.zeros(0, dtype=np.int64), max_cycle=1, thresh=0.0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        GR.minimize_orb_corr_GD(gamma, Gamma, inact, np.zeros(0, dtype=np.int64), max_cycle=4, thresh=0.0)
        torch.cuda.synchronize()
        phases['gd_newton_iteration_ms'] = (time.perf_counter() - t0) * 1e3 / 4
        # BASELINE configs[1..2] (method='nr' at n = 28 and the Cr2 shape n = 86): the whole call through the public API from HOST
        # arrays -- prep_rdm12 (H2D inside), four Newton-Raphson iterations, U / gamma / Gamma kept on the device, cost read back
        nr = {}
        for n_nr in (28, 86):
            d1_nr, d2_nr = synth.solver_rdms(n_nr)
            inact_nr = list(range(n_nr))

            def nr_call():
                g_, G_ = prep_rdm12(d1_nr, d2_nr, on_device=True)
                out_ = GR.minimize_orb_corr_GD(g_, G_, inact_nr, np.zeros(0, dtype=np.int64), max_cycle=4, thresh=0.0)
                torch.cuda.synchronize()
                return out_
            nr_call()
            t0 = time.perf_counter()
            nr_call()
            nr[str(n_nr)] = {'e2e_ms_per_call_4_iterations': (time.perf_counter() - t0) * 1e3, 'iterations': 4,
                             'h2d_bytes': int(8 * (n_nr ** 4 + n_nr ** 2))}
        phases['nr_method_e2e'] = nr

    acc_mask = rs['accepted'] != 0
    algo_bytes = 16.0 * (n ** 4 - (n - 2) ** 4) * n_accepted      # bytes the reference's update touches
    moved = moved_bytes(n, n, pairs_np, acc_mask)                 # what the deferred layout with dormant-row skipping moves
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    kname = ('sweep3_kernel (pipelined persistent Jacobi cycle: search CTA + reducers + column-owning workers)' if engine == 3
             else 'sweep_kernel (persistent Jacobi cycle, one grid barrier per pair)')
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'sweep_traffic.json')
    if os.path.exists(tpath):
        ent = json.load(open(tpath)).get(f'engine{engine}', {}).get(str(n))
        if ent:
            traffic = ent['dram_bytes_per_accepted_rotation'] * n_accepted
            traffic_src = f"{ent['source']}: {ent['dram_bytes_per_accepted_rotation'] / 1e6:.1f} MB per accepted rotation (ncu, {ent['accepted']} rotations) x {n_accepted}"
    roofline = {'kernel': kname,
                'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'peak_source': peak_src, 'traffic': traffic, 'traffic_source': traffic_src, 'algorithmic_bytes_per_launch': algo_bytes,
                'avg_launch_ms': sweep_ms, 'launches_timed': len(sweep_in_step),
                'moved_bytes_per_launch': moved,
                'frac_moved': moved / (sweep_ms * 1e-3) / 1e9 / peak,
                'frac_dram': (traffic / (sweep_ms * 1e-3) / 1e9 / peak) if traffic else None,
                'note': 'achieved / frac follow the contract: ALGORITHMIC bytes = 16 (n^4-(n-2)^4) per accepted rotation, the '
                        "reference's own touch count (SURVEY 8d) -- an effective bandwidth, it exceeds 1 when the layout moves fewer "
                        'bytes than the reference touches.  frac_moved = bytes the deferred-axis layout with dormant-row skipping really '
                        'reads + writes (64 n^2 n_active per rotation) / time / peak; frac_dram = DRAM bytes of the committed ncu capture '
                        '/ time / peak = the true HBM fraction.  Per-rotation working set %.0f MB vs 126 MB L2.' % (32.0 * n ** 3 / 1e6)}

    # the other two kernels north_star names: prep_rdm12 (HBM-bound) and the 4-index rotation (FP64 tensor pipe)
    prep_bytes = 16.0 * n ** 4
    roofline_prep = {'kernel': 'prep_rdm2_kernel', 'bound': 'hbm', 'achieved': prep_bytes / (phases['prep_rdm12_ms'] * 1e-3) / 1e9,
                     'peak': peak, 'unit': 'GB/s', 'frac': prep_bytes / (phases['prep_rdm12_ms'] * 1e-3) / 1e9 / peak,
                     'algorithmic_bytes_per_launch': prep_bytes, 'avg_launch_ms': phases['prep_rdm12_ms'], 'peak_source': peak_src}
    dmma_peak, dmma_src = measure_fp64_peak(timed)
    rot_flops = 8.0 * n ** 5
    rot_tf = rot_flops / (phases['rotate_rdm2_ms'] * 1e-3) / 1e12
    roofline_rotate = {'kernel': 'rotate_tma_kernel x 4 passes (4-index rotation, FP64 DMMA fed by TMA)', 'bound': 'tensor',
                       'achieved': rot_tf, 'peak': dmma_peak, 'unit': 'TFLOP/s', 'frac': rot_tf / dmma_peak, 'flops_per_launch_set': rot_flops,
                       'ms': phases['rotate_rdm2_ms'], 'peak_source': dmma_src}

    # ---- e2e through the public API from pinned host memory ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        api_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            api_step()      # all-pairs pass + sweep through the API; results come back to the host inside the call
        e2e_s = (time.perf_counter() - t0) / steps
        h2d = 8 * (n ** 4 + n ** 2) + 8 * n * n + 8 * (len(pairs_np) + P)
        d2h = LS_DTYPE.itemsize * (len(pairs_np) + P) + 8 * n * n + 8 * n * n + 32
        e2e = {'value': evals_per_step / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_s * 1e3,
               'api': 'qio_b200.qio.prep_rdm12(host dm1, pinned host dm2) -> grad.jacobi.linesearch_pairs -> mutual_info.mutual_information '
                      '-> grad.jacobi.minimize_orb_corr_jacobi(max_cycle=1, decision audit on); rotations, U, records, MI matrix read back'}

    cpu = None
    if detail and not args.no_cpu_baseline:
        cpu = cpu_reference_sample(n)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'blas_threads')}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(n, 1), 'clocks': clocks, 'e2e': e2e,
        'gpu_launches': launches_per_step * steps, 'roofline': roofline, 'roofline_prep': roofline_prep,
        'roofline_rotate': roofline_rotate, 'cpu_baseline': cpu,
        'evals_per_step': evals_per_step, 'rotations_accepted': n_accepted, 'cost_after_cycle': cost_after,
        'decision_audit': decision_audit, 'phases': phases, 'peak_memory_gb': peak_mem,
    }
    del dm2_d, Gamma, work, dm2_h
    return out


def measure_fp64_peak(timed):
    """In-run FP64 matrix peak for the rotation's roofline: cuBLAS DGEMM 8192^3 through torch (library call, used as a
    yardstick only; the hand-written DMMA microbenchmark of profiles/tools/fp64_peak.cu measured 37.1 TFLOP/s)."""
    import torch
    try:
        m = 8192
        a = torch.randn((m, m), dtype=torch.float64, device='cuda')
        b = torch.randn((m, m), dtype=torch.float64, device='cuda')
        torch.matmul(a, b)
        ms = min(timed(lambda: torch.matmul(a, b), 2) for _ in range(3))
        tf = 2.0 * m ** 3 / (ms * 1e-3) / 1e12
        del a, b
        return tf, f'measured in this run: cuBLAS DGEMM {m}^3 (torch.matmul f64), best of 3 x 2'
    except Exception as exc:       # noqa: BLE001
        return 37.1, f'fallback: DMMA m8n8k4 microbenchmark of round 1 (in-run DGEMM failed: {exc})'


def chain_table(trace):
    """Averages (us) along the per-pair dependency chain from the event trace of the timed sweep kernel (csrc/sweep3.cu TR_*):
    angle k-3 published -> worker sees it -> worker step starts -> partial of super block k posted -> reducer has all
    partials -> total published -> prefetch warps have it -> reduced super block ready -> search k starts -> angle k published.
    The sum is three pair periods (the lookahead is three pairs)."""
    tr = np.asarray(trace, dtype=np.float64).reshape(-1, 16)
    names = ['angle_to_worker', 'worker_plan_to_step_start', 'step_start_to_partial_posted', 'posted_to_reducer_has_all',
             'reducer_sum_publish', 'published_to_prefetch_has_it', 'prefetch_contract', 'ready_to_search_start', 'search']
    rows = []
    for k in range(3, tr.shape[0]):
        a, b = tr[k - 3], tr[k]
        ev = [a[1], a[2], a[3], b[4], b[5], b[6], b[7], b[8], b[0], b[1]]
        if min(ev) <= 0:
            continue
        rows.append(np.diff(np.array(ev)) / 1e3)
    if not rows:
        return None
    m = np.mean(np.array(rows), axis=0)
    out = {n_: float(v) for n_, v in zip(names, m)}
    out['three_pair_loop_us'] = float(m.sum())
    out['pairs_traced'] = len(rows)
    # the raw stamps of the first 12 traced pairs, us since the first one (columns: TR_* 0..9 of csrc/sweep3.cu)
    t0 = tr[tr > 0].min() if (tr > 0).any() else 0.0
    out['raw_us'] = [[round((v - t0) / 1e3, 2) if v > 0 else None for v in row[:13]] for row in tr[:12]]
    return out


def sweep_phase_table(tsum, npairs):
    """Per-pair averages (us) of the phase timers the sweep kernel keeps."""
    us = lambda v: float(v) / 1e3 / max(npairs, 1)
    if len(tsum) > 24 and float(tsum[24]) == 3.0:      # pipelined kernel (sweep3.cu): timers of the search CTA
        return {'engine': 3, 'cta0_wait_superblock': us(tsum[5]), 'cta0_setup_contract': us(tsum[9]), 'cta0_search': us(tsum[6]),
                'cta0_publish': us(tsum[7]), 'worker0_wait_angle_plan': us(tsum[16]), 'worker0_step': us(tsum[17]),
                'worker0_tile_phase': us(tsum[18]), 'worker0_pure_wait': us(tsum[19]), 'reducer0_wait': us(tsum[20]), 'reducer0_sum_publish': us(tsum[21]),
                'worker0_tile_pass': us(tsum[22]), 'worker0_flat_thread128': us(tsum[23]),
                'worker0_t0_until_dots': us(tsum[25]), 'worker0_t0_dots': us(tsum[26]), 'worker0_t0_sync': us(tsum[27]), 'worker0_t0_post': us(tsum[28])}
    return {'engine': 2, 'cta0_gather': us(tsum[5]), 'cta0_search': us(tsum[6]), 'cta0_tail': us(tsum[7]), 'cta0_barrier_wait': us(tsum[8]),
            'cta1_superblock_orbits': us(tsum[16]), 'cta1_flat_update': us(tsum[17]), 'cta1_barrier_wait': us(tsum[18]),
            'flatcta_update': us(tsum[22]), 'flatcta_barrier_wait': us(tsum[23]),
            'prologue_total_us': float(tsum[4]) / 1e3}


def run_b200_multi(args, rank, world, local_rank):
    """N > 1: rdm2 sharded by leading index, one process per GPU (qio_b200.sharded).  Same step as the 1-GPU arm."""
    import torch
    import torch.distributed as dist

    from qio_b200 import selfcheck
    from qio_b200.peers import PeerGroup

    sys.stderr.write(f'[bench] rank {rank}: NCCL communicator of {dist.get_world_size()} ranks (torch.distributed, backend nccl)\n')
    peers = PeerGroup()
    parity = None
    if not args.no_parity:
        # the sharded path against the single-GPU path on the very GPUs of this run, and the mailbox signalling itself
        par = selfcheck.multi_gpu_parity(peers, sizes=(13, 24))
        stress = selfcheck.mailbox_stress(peers, iterations=2000)
        parity = {'ok': bool(par['ok'] and stress['ok']), 'checks': par['checks'], 'failures': par['failures'][:5], 'sizes': par['sizes'],
                  'mailbox_stress': stress,
                  'what': 'qio_b200.selfcheck: sharded prep / rotation (all-to-all re-shard) / all-pairs pass / MI matrix / Jacobi cycles / flush vs '
                          'the single-GPU path on every rank, decisions bitwise identical; then tagged 16-byte stores into every peer mailbox'}
    out = multi_gpu_leg(args, args.n, args.steps, max(args.warmup, 3), rank, world, local_rank, peers, detail=True)
    if rank == 0:
        out['parity_ok'] = None if parity is None else parity['ok']
        out['parity'] = parity
        out['comm_nranks'] = world
    if not args.no_north_star and args.n != NORTH_STAR_N:
        torch.cuda.empty_cache()
        ns = multi_gpu_leg(args, NORTH_STAR_N, max(1, min(args.steps, 2)), 3, rank, world, local_rank, peers, detail=False)
        if rank == 0:
            out['north_star'] = north_star_record(ns)
    peers.close()
    dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(out), flush=True)


def multi_gpu_leg(args, n, steps, warmup, rank, world, local_rank, peers, detail):
    import torch
    import torch.distributed as dist

    from qio_b200 import device as dev
    from qio_b200 import synth, tables
    from qio_b200._lib import LS_DTYPE
    from qio_b200.grad import jacobi as JA
    from qio_b200.sharded import ShardedRDM, shard_range

    P = n * (n - 1) // 2
    inact = list(range(n))
    a0, a1 = shard_range(n, rank, world)
    na = a1 - a0
    # this rank's slabs of the synthetic solver output, on the device and in pinned host memory
    w, Pk = synth.projectors(n)
    d1 = 2.0 * torch.from_numpy(Pk).cuda()
    wt = torch.from_numpy(w).cuda()
    dm1_d = torch.einsum('k,kab->ab', wt, d1).contiguous()
    dm2_d = torch.zeros((na, n, n, n), dtype=torch.float64, device='cuda')
    for k in range(len(w)):     # input synthesis only (plumbing, outside every timed region)
        dl = d1[k][a0:a1]
        dm2_d += wt[k] * (torch.einsum('pq,rs->pqrs', dl, d1[k]) - 0.5 * torch.einsum('ps,rq->pqrs', dl, d1[k]))
    dm2_h = torch.empty((na, n, n, n), dtype=torch.float64).pin_memory()
    dm2_h.copy_(dm2_d)
    dm1_h = dm1_d.cpu().numpy()
    torch.cuda.synchronize()

    U0, order = sweep_order(n)
    pairs_np = JA.sweep_pairs(order, inact)
    U0_d = dev.to_device(U0)
    mask = dev.inactive_mask(inact, n)
    inds_d = dev.to_device(np.arange(n, dtype=np.int32), dev.I32)
    tabs = tables.device_tables()
    fine_len = tabs.host['fine_len']
    W = dev.empty(n, n)
    state = {}
    torch.cuda.reset_peak_memory_stats()
    mem_before = torch.cuda.memory_allocated()

    sweep_events = []           # CUDA events around the sweep kernel of every step

    def step(dm1_in, dm2_in):
        state.pop('st', None)       # release the previous step's shard buffers before the new ones are allocated
        st = ShardedRDM.from_solver(dm1_in, dm2_in, n, peers=peers)
        st.rotate_(U0_d)
        dX, ddX, rec_all = st.all_pairs_pass(inact, tabs)
        mi = st.mutual_information()
        Ud = U0_d.clone()
        dev.sweep_reset(W, n)
        rec, n_acc, cost, flags = st.sweep_cycle(Ud, W, pairs_np, mask, inds_d, tabs)      # includes the decision audit
        sweep_events.append(st.last_sweep_events)
        st.flush(W)
        state.update(rec_all=rec_all, rec=rec, n_acc=n_acc, cost=cost, U=Ud, st=st, mi=mi, audit=st.last_audit)

    for _ in range(warmup):
        step(dm1_d, dm2_d)
    torch.cuda.synchronize()
    peak_mem = (torch.cuda.max_memory_allocated() - mem_before) / 1e9
    ra = state['rec_all'].cpu().numpy().view(LS_DTYPE)
    rs = state['rec']
    rows_a = np.where(ra['coarse_index'] >= 0, ra['coarse_index'], 0)
    rows_s = np.where(rs['coarse_index'] >= 0, rs['coarse_index'], 0)
    evals_per_step = int((1 + 314 + fine_len[rows_a]).sum() + (1 + 314 + fine_len[rows_s]).sum())
    audit = state['audit']
    decision_audit = {'pairs': audit['pairs'], 'mismatches': len(audit['mismatches']), 'near_ties_below_1e-12': len(audit['near_ties']),
                      'unexplained': len(audit['unexplained']), 'min_fine_margin': audit['min_fine_margin'],
                      'min_coarse_margin': audit['min_coarse_margin']}

    step_ms_each = []

    def timed_steps(dm1_in, dm2_in, k):
        sampler = ClockSampler(local_rank)
        sampler.start()
        sampler.wait_ready()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        e0.record()
        marks[0].record()
        for q in range(k):
            step(dm1_in, dm2_in)
            marks[q + 1].record()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)] + [marks[q].elapsed_time(marks[q + 1]) for q in range(k)], dtype=torch.float64, device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        step_ms_each.append([float(x) for x in ms[1:].tolist()])        # diagnostics: every step of the timed region, max over ranks
        return float(ms[0].item()) / k, sampler.stop()

    del sweep_events[:]
    ms_per_step, clocks = timed_steps(dm1_d, dm2_d, steps)
    value = evals_per_step / (ms_per_step * 1e-3)
    # the dominant kernel inside the timed region: per launch the slowest rank, then the average over the launches
    sw = torch.tensor([a.elapsed_time(b) for a, b in sweep_events], dtype=torch.float64, device='cuda')
    dist.all_reduce(sw, op=dist.ReduceOp.MAX)
    sweep_in_step = [float(x) for x in sw.tolist()]
    e2e = None
    if not args.no_e2e:
        step(dm1_h, dm2_h)
        e2e_ms, _ = timed_steps(dm1_h, dm2_h, steps)
        e2e = {'value': evals_per_step / (e2e_ms * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(8 * (n ** 4 + n ** 2) + 8 * (len(pairs_np) + P)),
               'd2h_bytes_per_step': int(LS_DTYPE.itemsize * (len(pairs_np) + P) + 64 * world), 'ms_per_step': e2e_ms,
               'api': 'qio_b200.sharded.ShardedRDM.from_solver(host dm1, pinned host dm2 shard) -> rotate_ -> all_pairs_pass '
                      '-> mutual_information -> sweep_cycle (decision audit on) -> flush on every rank; records read back'}

    # ---- phases: the sweep kernel alone (dominant kernel) and what surrounds it, from the step's own start state -----------
    def tmax(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), r

    phases = {}
    phases['prep_upload_ms'], st = tmax(lambda: ShardedRDM.from_solver(dm1_d, dm2_d, n, peers=peers))
    phases['rotate_sharded_ms'], _ = tmax(lambda: st.rotate_(U0_d))
    phases['all_pairs_pass_ms'], _ = tmax(lambda: st.all_pairs_pass(inact, tabs))
    phases['mutual_information_ms'], _ = tmax(lambda: st.mutual_information())
    Ud = U0_d.clone()
    dev.sweep_reset(W, n)
    pairs_d = dev.to_device(pairs_np, dev.I32)
    alone_ms, (_, summ) = tmax(lambda: dev.sweep_cycle(st.gamma, st.G, Ud, W, n, pairs_d, mask, tabs, st.a0, st.na, peers))
    n_acc = int(dev.to_host(summ)[0])
    phases['jacobi_cycle_kernel_alone_ms'] = alone_ms      # one extra launch after a barrier: launch skew of the ranks is inside
    sweep_ms = float(np.mean(sweep_in_step))               # average launch duration over the timed steps (max over ranks each)
    phases['jacobi_cycle_kernel_ms'] = sweep_ms
    phases['jacobi_cycle_kernel_ms_each'] = sweep_in_step
    phases['step_ms_each'] = step_ms_each[0] if step_ms_each else None
    phases['flush_sharded_ms'], _ = tmax(lambda: st.flush(W))
    phases['step_minus_sweep_kernel_ms'] = ms_per_step - sweep_ms
    if detail:
        trace = dev.empty(32, 16)
        tsum = dev.to_host(dev.sweep_cycle(st.gamma, st.G, Ud, W, n, pairs_d, mask, tabs, st.a0, st.na, peers, timers=True,
                                           trace_out=trace, trace_first_pair=len(pairs_np) // 2)[1])
        phases['sweep_phase_us_per_pair'] = sweep_phase_table(tsum, len(pairs_np))
        phases['sweep_chain_us'] = chain_table(dev.to_host(trace))
    peak, peak_src = hbm_peak()
    acc_mask = rs['accepted'] != 0
    algo_bytes = 16.0 * (n ** 4 - (n - 2) ** 4) * n_acc / world
    moved = moved_bytes(n, na, pairs_np, acc_mask)
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    engine = dev.sweep_engine(n, na)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'sweep_traffic.json')
    if os.path.exists(tpath):
        ent = json.load(open(tpath)).get(f'engine{engine}', {}).get(f'{n}x{world}')
        if ent:
            traffic = ent['dram_bytes_per_accepted_rotation'] * n_acc
            traffic_src = f"{ent['source']} (per rank)"
    roofline = {'kernel': ('sweep3_kernel (pipelined persistent Jacobi cycle)' if engine == 3 else 'sweep_kernel (persistent Jacobi cycle, grid barrier per pair)') + ', per GPU', 'bound': 'hbm', 'achieved': achieved,
                'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                'algorithmic_bytes_per_launch': algo_bytes, 'avg_launch_ms': sweep_ms, 'launches_timed': len(sweep_in_step),
                'moved_bytes_per_launch': moved, 'frac_moved': moved / (sweep_ms * 1e-3) / 1e9 / peak,
                'frac_dram': (traffic / (sweep_ms * 1e-3) / 1e9 / peak) if traffic else None,
                'peak_source': peak_src,
                'note': 'per-GPU share of the algorithmic bytes (16 (n^4-(n-2)^4) per accepted rotation / n_gpus) over the '
                        'slowest rank\'s kernel time; frac_moved = bytes this rank\'s shard really reads + writes / time / peak'}
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
        'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(n, world),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': count_launches(engine, world) * steps,
        'roofline': roofline, 'cpu_baseline': None, 'evals_per_step': evals_per_step,
        'rotations_accepted': n_acc, 'cost_after_cycle': state['cost'], 'decision_audit': decision_audit,
        'phases': phases, 'peak_memory_gb': peak_mem,
    }
    del dm2_d, dm2_h, st
    state.clear()
    return out


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
