"""One sweep_cycle launch (n orbitals, first P pairs of the seeded sweep order) for ncu captures of sweep_kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qio_b200 import device as dev  # noqa: E402
from qio_b200 import synth, tables  # noqa: E402
from qio_b200.grad import jacobi as JA  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
P = int(sys.argv[2]) if len(sys.argv) > 2 else 600
gamma, Gamma = synth.prepared_rdms(n)
np.random.seed(0)
U0 = JA.random_start(n)
order = np.arange(n)
np.random.shuffle(order)
inact = list(range(n))
pairs = JA.sweep_pairs(order, inact)[:P]
g, G, Ud = dev.to_device(gamma), dev.to_device(Gamma), dev.to_device(U0)
g = dev.rotate_rdm1(Ud, g, n)
dev.rotate_rdm2_(Ud, G, n)
W = dev.empty(n, n)
dev.sweep_reset(W, n)
res, summ = dev.sweep_cycle(g, G, Ud.clone(), W, n, dev.to_device(pairs, dev.I32), dev.inactive_mask(inact, n), tables.device_tables())
torch.cuda.synchronize()
print('accepted', int(summ[0].item()), 'of', len(pairs))
