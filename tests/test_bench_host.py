"""CPU tests of bench.py's host side: the byte model behind roofline.frac_moved, the chain table of the event trace, the sweep
order, and the JSON contract of the reference arm (run on a tiny problem; it must never map the CUDA library)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_moved_bytes_matches_a_direct_count():
    """moved_bytes = sum over accepted rotations of 2 axes x 2 rows x (orbitals awake at that step) x n x na x 8 B, read + written;
    an orbital is awake from three pairs before its first appearance in the pair list."""
    n, na, look = 7, 3, 3
    pairs = np.array([[0, 1], [2, 3], [0, 2], [4, 5], [1, 6], [3, 4]], dtype=np.int32)
    accepted = np.array([1, 0, 1, 1, 0, 1], dtype=bool)
    total = 0
    for w in range(len(pairs)):
        if not accepted[w]:
            continue
        awake = {int(x) for k in range(min(len(pairs), w + look + 1)) for x in pairs[k]}
        total += 2 * 2 * len(awake) * n * na * 8 * 2
    assert bench.moved_bytes(n, na, pairs, accepted, look) == float(total)
    # nothing accepted: nothing moved; everything accepted on a fully awake set: the closed form
    assert bench.moved_bytes(n, na, pairs, np.zeros(6, dtype=bool)) == 0.0
    full = np.array([[i, j] for i in range(4) for j in range(i)], dtype=np.int32)
    assert bench.moved_bytes(4, 4, full, np.ones(len(full), dtype=bool), look=100) == len(full) * 64.0 * 4 * 4 * 4


def test_chain_table_segments_add_up_to_the_three_pair_loop():
    """Synthetic stamps: every event of pair k at a known offset; the table returns the differences along the chain
    (pair k-3's angle -> ... -> pair k's angle) in microseconds and their sum as the loop time."""
    P = 10
    tr = np.zeros((32, 16))
    period = 7000.0                          # ns between pairs
    off = {0: 9000.0, 1: 13000.0, 2: 14000.0, 3: 15500.0, 4: 1000.0, 5: 4000.0, 6: 4500.0, 7: 6500.0, 8: 8800.0}
    for k in range(P):
        for ev, o in off.items():
            tr[k, ev] = 1e9 + k * period + o
    tab = bench.chain_table(tr)
    assert tab['pairs_traced'] == P - 3
    want = [14000 - 13000, 15500 - 14000, 3 * period + 1000 - 15500, 4000 - 1000, 4500 - 4000, 6500 - 4500, 8800 - 6500, 9000 - 8800,
            13000 - 9000]
    names = ['angle_to_worker', 'worker_plan_to_step_start', 'step_start_to_partial_posted', 'posted_to_reducer_has_all',
             'reducer_sum_publish', 'published_to_prefetch_has_it', 'prefetch_contract', 'ready_to_search_start', 'search']
    for nm, w in zip(names, want):
        assert abs(tab[nm] - w / 1e3) < 1e-9, nm
    assert abs(tab['three_pair_loop_us'] - 3 * period / 1e3) < 1e-9
    assert bench.chain_table(np.zeros((32, 16))) is None


def test_sweep_order_is_the_reference_rng_sequence():
    """The bench workload's start rotation and orbital order come from numpy's legacy global RNG seeded with 0, in the reference's
    call order (qio/grad/jacobi.py:183-203)."""
    U1, o1 = bench.sweep_order(12)
    U2, o2 = bench.sweep_order(12)
    assert np.array_equal(U1, U2) and np.array_equal(o1, o2)
    assert sorted(o1.tolist()) == list(range(12))
    assert np.allclose(U1 @ U1.T, np.eye(12), atol=1e-12)


def test_reference_arm_json_contract_and_no_product_library():
    """`bench.py --impl reference` prints ONE JSON line with the contract's keys, runs the reference's own functions
    (kind 'reference' when oracle/_ref exists, else the port) and never maps libqio_b200.so."""
    env = dict(os.environ)
    env.pop('RANK', None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--orbitals', '10', '--steps', '1',
                          '--warmup', '0'], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in rec, key
    assert rec['impl'] == 'reference' and rec['value'] > 0 and rec['gpu_launches'] == 0
    assert rec['loaded_product_library'] is False
    cb = rec['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == rec['value']
    assert rec['e2e'] == {'value': rec['value'], 'unit': rec['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    if os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'qio')):
        assert cb['kind'] == 'reference'
