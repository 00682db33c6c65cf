#!/usr/bin/env python
"""Benchmark of the QIO orbital-entropy hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 100] [--impl reference]

One "step" is one complete pass of the hot path over one synthetic solver output:

    prep_rdm12(dm1, dm2)                        a-1   (restores the pristine RDMs every step)
    random-start orbital rotation (2- and 4-index, FP64 DMMA)   a-7 start / a-10 rotation
    all-pairs pass: pair blocks, gradient + Hessian, angle line search for all n(n-1)/2 pairs   a-8, a-5
    one full Jacobi cycle: 4950 (n=100) sequential line searches + Givens updates of the RDMs   a-7, a-6
    get_cost_fqi of the result                   a-3

metric  = orbital-pair entropy evaluations (jacobi_cost calls of the reference) per second over the step;
ms_per_step is the wall time of the whole step (all-pairs pass + Jacobi sweep).
`value`: inputs (dm1, dm2, start rotation, pair order) resident in HBM.   `e2e`: the same step through the
public Python API from pinned HOST buffers, H2D of dm1/dm2 and D2H of the results inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'orbital_pair_entropy_evals_per_s'
UNIT = 'evals/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    # '--orbitals' is the spelling to use under torchrun: its own parser rejects '--n' as an ambiguous prefix
    ap.add_argument('--n', '--orbitals', dest='n', type=int, default=int(os.environ.get('QIO_BENCH_N', '100')),
                    help='number of orbitals (100 = configs[3], 200 = configs[4])')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    return ap.parse_args()


# --------------------------------------------------------------------------------------------------
# workload description shared by both arms
# --------------------------------------------------------------------------------------------------
def workload_config(n, n_gpus):
    return {
        'workload': f'synthetic determinant-ensemble rdm1/rdm2, n={n} orbitals (K=4, nocc={max(1, n // 5)}, seed 20261019+n), '
                    f'inactive = all orbitals: prep_rdm12 + random-start rotation + all-pairs grad/Hess/line-search/pair-entropy '
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
# reference arm / cpu baseline: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
_CPU_SETUP = {}


def cpu_reference_setup(n):
    """Rotated synthetic RDMs and the sweep pair order on the host (built once per process)."""
    if n not in _CPU_SETUP:
        from oracle import qio_oracle as O
        from qio_b200 import synth
        t0 = time.time()
        gamma0, Gamma0 = synth.prepared_rdms(n)
        U0, order = sweep_order(n)
        g = O.rotate_rdm1(U0, gamma0)
        G = O.rotate_rdm2(U0, Gamma0)
        inact = list(range(n))
        _CPU_SETUP[n] = dict(g=g, G=G, inact=inact, pairs=O.sweep_pairs(order, inact), setup_s=time.time() - t0)
    return _CPU_SETUP[n]


def cpu_reference_sample(n, n_ls_pairs=48, n_sweep_pairs=4, offset=0):
    """Times the reference-shaped CPU path (oracle port: scalar jacobi_cost loop for the line search, dense
    einsum for jacobi_transform -- what qio/grad/jacobi.py executes) on a bounded sample and extrapolates the
    evals/s of the full step:  all-pairs pass = P line searches;  sweep = P x (line search + accepted transform)."""
    from oracle import qio_oracle as O
    st = cpu_reference_setup(n)
    g, G, inact, pairs = st['g'], st['G'], st['inact'], st['pairs']
    P = len(pairs)
    sample = [pairs[(offset + k) % P] for k in range(n_ls_pairs)]
    # line searches (1 core, pure Python like the reference)
    t0 = time.time()
    thetas = [O.jacobi_direct_scalar(i, j, g, G, inact) for (i, j) in sample]
    t_ls = (time.time() - t0) / n_ls_pairs
    accepted = sum(t is not None for t in thetas)
    # dense transforms (BLAS threads)
    t0 = time.time()
    done = 0
    for (i, j), t in zip(sample[:n_sweep_pairs], thetas[:n_sweep_pairs]):
        O.jacobi_transform_dense(g, G, i, j, 0.1 if t is None else t)
        done += 1
    t_tr = (time.time() - t0) / max(done, 1)
    acc_rate = accepted / n_ls_pairs
    evals_per_pair = 515.5
    step_s = P * t_ls + P * (t_ls + acc_rate * t_tr)
    value = 2 * P * evals_per_pair / step_s
    return {
        'value': value, 'unit': UNIT, 'cores': os.cpu_count(), 'kind': 'port',
        'sample': (f'oracle port of the reference CPU path at n={n}: {n_ls_pairs} scalar line searches '
                   f'({t_ls * 1e3:.1f} ms each, 1 core) + {done} dense jacobi_transform einsums ({t_tr:.2f} s each, '
                   f'BLAS on {os.cpu_count()} cores); full step extrapolated as P*t_ls + P*(t_ls + {acc_rate:.2f}*t_tr) '
                   f'= {step_s:.0f} s for P={P} pairs'),
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
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': last['cores'], 'kind': 'port', 'sample': last['sample']},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    if not os.environ.get('QIO_BENCH_KEEP_NCCL_DEBUG'):
        os.environ['NCCL_DEBUG'] = 'WARN'       # keep NCCL's version banner off stdout: rank 0 prints one JSON line only
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    if args.gpus != world and world > 1:
        args.gpus = world

    from qio_b200 import device as dev
    from qio_b200 import synth, tables
    from qio_b200._lib import LS_DTYPE
    from qio_b200.grad import jacobi as JA
    from qio_b200.qio import prep_rdm12

    if world > 1:
        return run_b200_multi(args, rank, world, local_rank)

    n = args.n
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
    for k in range(len(w)):     # input synthesis only (plumbing, outside every timed region)
        dm2_d += wt[k] * (torch.einsum('pq,rs->pqrs', d1[k], d1[k]) - 0.5 * torch.einsum('ps,rq->pqrs', d1[k], d1[k]))
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
    inds_d = dev.to_device(np.arange(n, dtype=np.int32), dev.I32)
    state = {}

    def device_step():
        """The whole hot path with every input already in HBM."""
        dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma)
        g = dev.rotate_rdm1(U0_d, gamma, n)
        dev.rotate_rdm2_(U0_d, Gamma, n, work)
        blocks = dev.pair_blocks(g, Gamma, n, allpairs_d)
        dX, ddX = dev.fqi_grad_hess(blocks, n, mask)
        ls_all = dev.pair_linesearch(blocks, allpairs_d, mask, tabs)
        s_pair, _, _ = dev.pair_entropy(blocks)          # extension a-9: pair entropies for the mutual-information matrix
        Ud = U0_d.clone()
        dev.sweep_reset(W, n)
        res, summary = dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs)
        diag = dev.deferred_diagonals(g, Gamma, W, n, inds_d)
        _, _, _, esum = dev.orbital_entropies(diag, want_spec=False)
        dev.zero_offspin_if_rotated(g, n, summary)
        dev.sweep_flush(Gamma, W, n, work)
        state.update(ls_all=ls_all, res=res, summary=summary, esum=esum, U=Ud, dX=dX, g=g)

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
        np.random.seed(0)
        rots, U, g1, G1 = JA.minimize_orb_corr_jacobi(g0, G0, inact, 1)
        torch.cuda.synchronize()
        return rec_all, rots, U

    # ---- warm-up ------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        device_step()
    torch.cuda.synchronize()
    evals_per_step, n_accepted, rs = count_evals()
    cost_after = float(dev.to_host(state['esum'])[0])

    # ---- timed: device-resident -----------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    ms_per_step = ms_total / args.steps
    value = evals_per_step / (ms_per_step * 1e-3)
    launches_per_step = 2 + 1 + 4 + 1 + 1 + 1 + 1 + (1 + 1 + 1 + 1 + 1) + 3   # prep, rotations, all-pairs (+MI), sweep (+cost), flush
    if dev.sweep_engine(n) == 3:
        launches_per_step += 1      # per-pair table kernel of the pipelined sweep

    # ---- phase timings (explain the step) and the roofline of the dominant kernel ----------------------
    def timed(fn, reps=1):
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a_.record()
        for _ in range(reps):
            fn()
        b_.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / reps

    phases = {}
    phases['prep_rdm12_ms'] = timed(lambda: dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma), 3)
    phases['rotate_rdm2_ms'] = timed(lambda: dev.rotate_rdm2_(U0_d, Gamma, n, work), 3)
    g = dev.rotate_rdm1(U0_d, gamma, n)
    blocks = dev.pair_blocks(g, Gamma, n, allpairs_d)
    phases['pair_blocks_ms'] = timed(lambda: dev.pair_blocks(g, Gamma, n, allpairs_d), 5)
    phases['grad_hess_ms'] = timed(lambda: dev.fqi_grad_hess(blocks, n, mask), 5)
    phases['linesearch_all_pairs_ms'] = timed(lambda: dev.pair_linesearch(blocks, allpairs_d, mask, tabs), 5)
    phases['pair_entropy_mi_ms'] = timed(lambda: dev.pair_entropy(blocks), 5)
    Ud = U0_d.clone()
    dev.sweep_reset(W, n)
    sweep_ms = timed(lambda: dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs), 1)
    phases['jacobi_cycle_kernel_ms'] = sweep_ms
    rs2 = dev.results_to_numpy(dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs)[0])   # 3rd cycle: count below
    phases['sweep_flush_ms'] = timed(lambda: dev.sweep_flush(Gamma, W, n, work), 1)
    m1_evals = int((1 + 314 + fine_len[np.where(dev.results_to_numpy(state['ls_all'])['coarse_index'] >= 0,
                                                dev.results_to_numpy(state['ls_all'])['coarse_index'], 0)]).sum())
    phases['all_pairs_evals_per_s'] = m1_evals / (phases['linesearch_all_pairs_ms'] * 1e-3)

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

    # dominant kernel = the persistent sweep kernel.  Timed alone on the step's own first cycle (same start state).
    dev.prep_rdm12(dm1_d, dm2_d, n, gamma=gamma, Gamma=Gamma)
    g = dev.rotate_rdm1(U0_d, gamma, n)
    dev.rotate_rdm2_(U0_d, Gamma, n, work)
    Ud = U0_d.clone()
    dev.sweep_reset(W, n)
    sweep_ms = timed(lambda: dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs), 1)
    phases['jacobi_cycle_kernel_first_cycle_ms'] = sweep_ms
    os.environ['QIO_B200_SWEEP_TIMERS'] = '1'       # the build of the sweep kernel that carries its phase timers (about 5 % slower)
    tsum = dev.to_host(dev.sweep_cycle(g, Gamma, Ud, W, n, pairs_d, mask, tabs)[1])
    os.environ.pop('QIO_B200_SWEEP_TIMERS', None)
    phases['sweep_phase_us_per_pair'] = sweep_phase_table(tsum, len(pairs_np))
    lsn = ['coarse', 'reduce', 'fine', 'scan', 'margin', 'tail']
    phases['linesearch_cycles_per_pair_cta0'] = {k: float(v) / len(pairs_np) for k, v in zip(lsn, tsum[10:16])}
    algo_bytes = 16.0 * (n ** 4 - (n - 2) ** 4) * n_accepted      # bytes the reference update touches
    kernel_bytes = 64.0 * n ** 3 * n_accepted + 32.0 * n * n * len(pairs_np)   # what the deferred layout moves
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))['hbm_gbs']
        peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)'
    else:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    engine = dev.sweep_engine(n)
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
                'avg_launch_ms': sweep_ms, 'launches_timed': 1,
                'kernel_min_bytes_per_launch': kernel_bytes,
                'kernel_min_bytes_frac': kernel_bytes / (sweep_ms * 1e-3) / 1e9 / peak,
                'note': 'algorithmic bytes = 16 (n^4-(n-2)^4) per accepted rotation (SURVEY 8d); the deferred-axis layout '
                        'moves only 64 n^3 B per rotation (+ 32 n^2 B of block rows per pair), see kernel_min_bytes_*; '
                        'per-rotation working set %.0f MB vs 126 MB L2 (L2-resident: DRAM traffic below the algorithmic bytes, '
                        'frac above 1 is possible)' % (32.0 * n ** 3 / 1e6)}

    # ---- e2e through the public API from pinned host memory ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        api_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            api_step()      # all-pairs pass + sweep through the API; results come back to the host inside the call
        e2e_s = (time.perf_counter() - t0) / args.steps
        h2d = 8 * (n ** 4 + n ** 2) + 8 * n * n + 8 * (len(pairs_np) + P)
        d2h = LS_DTYPE.itemsize * (len(pairs_np) + P) + 8 * n * n + 32
        e2e = {'value': evals_per_step / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_s * 1e3,
               'api': 'qio_b200.qio.prep_rdm12(host dm1, pinned host dm2) -> grad.jacobi.linesearch_pairs -> '
                      'grad.jacobi.minimize_orb_corr_jacobi(max_cycle=1); rotations, U, records read back'}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_reference_sample(n)
        cpu = {k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(n, 1), 'clocks': clocks, 'e2e': e2e,
        'gpu_launches': launches_per_step * args.steps, 'roofline': roofline, 'cpu_baseline': cpu,
        'evals_per_step': evals_per_step, 'rotations_accepted': n_accepted, 'cost_after_cycle': cost_after,
        'phases': phases,
    }
    print(json.dumps(out))


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

    from qio_b200 import device as dev
    from qio_b200 import synth, tables
    from qio_b200._lib import LS_DTYPE
    from qio_b200.grad import jacobi as JA
    from qio_b200.peers import PeerGroup
    from qio_b200.sharded import ShardedRDM, shard_range

    n = args.n
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
    peers = PeerGroup()
    W = dev.empty(n, n)
    work_full = dev.empty(n, n, n, n)
    state = {}

    def step(dm1_in, dm2_in):
        st = ShardedRDM.from_solver(dm1_in, dm2_in, n, peers=peers)
        st.rotate_(U0_d, work_full)
        dX, ddX, rec_all = st.all_pairs_pass(inact, tabs)
        Ud = U0_d.clone()
        dev.sweep_reset(W, n)
        rec, n_acc, cost, flags = st.sweep_cycle(Ud, W, pairs_np, mask, inds_d, tabs)
        st.flush(W, work_full)
        state.update(rec_all=rec_all, rec=rec, n_acc=n_acc, cost=cost, U=Ud, st=st)

    for _ in range(max(args.warmup, 3)):
        step(dm1_d, dm2_d)
    torch.cuda.synchronize()
    ra = state['rec_all'].cpu().numpy().view(LS_DTYPE)
    rs = state['rec']
    rows_a = np.where(ra['coarse_index'] >= 0, ra['coarse_index'], 0)
    rows_s = np.where(rs['coarse_index'] >= 0, rs['coarse_index'], 0)
    evals_per_step = int((1 + 314 + fine_len[rows_a]).sum() + (1 + 314 + fine_len[rows_s]).sum())

    def timed_steps(dm1_in, dm2_in, k):
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(k):
            step(dm1_in, dm2_in)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / k, sampler.stop()

    ms_per_step, clocks = timed_steps(dm1_d, dm2_d, args.steps)
    value = evals_per_step / (ms_per_step * 1e-3)
    e2e = None
    if not args.no_e2e:
        step(dm1_h, dm2_h)
        e2e_ms, _ = timed_steps(dm1_h, dm2_h, args.steps)
        e2e = {'value': evals_per_step / (e2e_ms * 1e-3), 'unit': UNIT,
               'h2d_bytes_per_step': int(8 * (n ** 4 + n ** 2) + 8 * (len(pairs_np) + P)),
               'd2h_bytes_per_step': int(LS_DTYPE.itemsize * (len(pairs_np) + P) + 64 * world), 'ms_per_step': e2e_ms,
               'api': 'qio_b200.sharded.ShardedRDM.from_solver(host dm1, pinned host dm2 shard) -> rotate_ -> all_pairs_pass '
                      '-> sweep_cycle -> flush on every rank; records read back'}

    # the sweep kernel alone (dominant kernel), from the step's own start state
    st = ShardedRDM.from_solver(dm1_d, dm2_d, n, peers=peers)
    st.rotate_(U0_d, work_full)
    Ud = U0_d.clone()
    dev.sweep_reset(W, n)
    pairs_d = dev.to_device(pairs_np, dev.I32)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, summ = dev.sweep_cycle(st.gamma, st.G, Ud, W, n, pairs_d, mask, tabs, st.a0, st.na, peers)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sweep_ms = float(t.item())
    n_acc = int(dev.to_host(summ)[0])
    os.environ['QIO_B200_SWEEP_TIMERS'] = '1'       # one more cycle with the timed build, for the phase table only
    tsum = dev.to_host(dev.sweep_cycle(st.gamma, st.G, Ud, W, n, pairs_d, mask, tabs, st.a0, st.na, peers)[1])
    os.environ.pop('QIO_B200_SWEEP_TIMERS', None)
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peak = json.load(open(peaks_path))['hbm_gbs'] if os.path.exists(peaks_path) else 6650.0
    algo_bytes = 16.0 * (n ** 4 - (n - 2) ** 4) * n_acc / world
    achieved = algo_bytes / (sweep_ms * 1e-3) / 1e9
    engine = dev.sweep_engine(n, na)
    roofline = {'kernel': ('sweep3_kernel (pipelined persistent Jacobi cycle)' if engine == 3 else 'sweep_kernel (persistent Jacobi cycle, grid barrier per pair)') + ', per GPU', 'bound': 'hbm', 'achieved': achieved,
                'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'algorithmic_bytes_per_launch': algo_bytes, 'avg_launch_ms': sweep_ms, 'launches_timed': 1,
                'kernel_min_bytes_per_launch': (64.0 * n ** 3 * n_acc + 32.0 * n * n * len(pairs_np)) / world,
                'peak_source': 'measured (MEASURED_PEAKS.json hbm_gbs)' if os.path.exists(peaks_path) else 'fallback',
                'note': 'per-GPU share of the algorithmic bytes (16 (n^4-(n-2)^4) per accepted rotation / n_gpus) over the '
                        'slowest rank\'s kernel time'}
    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(n, world),
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': (2 + 1 + 3 * na + 1 + 1 + 2 + 5 + 3) * args.steps,
            'roofline': roofline, 'cpu_baseline': None, 'evals_per_step': evals_per_step,
            'rotations_accepted': n_acc, 'cost_after_cycle': state['cost'],
            'phases': {'jacobi_cycle_kernel_ms': sweep_ms,
                       'sweep_phase_us_per_pair': sweep_phase_table(tsum, len(pairs_np))},
        }
        print(json.dumps(out))
    peers.close()
    dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
