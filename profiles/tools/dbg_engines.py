"""Debug: first pair where the pipelined and the barrier sweep kernel disagree."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qio_b200 import device as dev, synth, tables
from qio_b200.grad import jacobi as JA
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
gamma, Gamma = synth.prepared_rdms(n)
np.random.seed(0)
U0 = JA.random_start(n)
order = np.arange(n); np.random.shuffle(order)
inact = list(range(n))
pairs = JA.sweep_pairs(order, inact)
print('order', order.tolist())
out = {}
for eng in ('3', '2'):
    os.environ['QIO_B200_SWEEP_ENGINE'] = eng
    g, G, Ud = dev.to_device(gamma), dev.to_device(Gamma), dev.to_device(U0)
    g = dev.rotate_rdm1(Ud, g, n); dev.rotate_rdm2_(Ud, G, n)
    W = dev.empty(n, n); dev.sweep_reset(W, n)
    res, summ = dev.sweep_cycle(g, G, Ud.clone(), W, n, dev.to_device(pairs, dev.I32), dev.inactive_mask(inact, n), tables.device_tables())
    torch.cuda.synchronize()
    out[eng] = (dev.results_to_numpy(res), G.cpu().numpy().copy(), W.cpu().numpy().copy())
a, b = out['3'][0], out['2'][0]
bad = [k for k in range(len(pairs)) if not (a['accepted'][k] == b['accepted'][k] and (a['theta'][k] == b['theta'][k] or (np.isnan(a['theta'][k]) and np.isnan(b['theta'][k]))))]
print('pairs', len(pairs), 'first mismatches', bad[:5])
for k in bad[:3]: print(k, pairs[k].tolist(), a['theta'][k], b['theta'][k], a['cost0'][k], b['cost0'][k])
k0 = bad[0] if bad else len(pairs)
print('cost0 diffs up to first mismatch:', [float(abs(a['cost0'][k] - b['cost0'][k])) for k in range(min(k0 + 1, len(pairs)))][:40])
print('S diff', float(np.abs(out['3'][1] - out['2'][1]).max()))
