"""One sweep_cycle launch (n orbitals, first P pairs of the seeded sweep order) for ncu captures of the sweep kernel.
The synthetic RDMs are built on the device (n = 200 would need 25 GB of host arrays otherwise).
    python profiles/tools/sweep_profile.py [n] [pairs] [engine]"""
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
engine = int(sys.argv[3]) if len(sys.argv) > 3 else 0
w, Pk = synth.projectors(n)
Pd = torch.from_numpy(Pk).cuda()
wt = torch.from_numpy(w).cuda()
g1 = torch.einsum('k,kab->ab', wt, Pd)
g = torch.zeros((2 * n, 2 * n), dtype=torch.float64, device='cuda')
g[0::2, 0::2] = g1
g[1::2, 1::2] = g1
G = torch.zeros((n, n, n, n), dtype=torch.float64, device='cuda')
for k in range(len(w)):
    for a0 in range(0, n, 25):
        G[a0:a0 + 25] += wt[k] * torch.einsum('ad,bc->abcd', Pd[k][a0:a0 + 25], Pd[k])
np.random.seed(0)
U0 = JA.random_start(n)
order = np.arange(n)
np.random.shuffle(order)
inact = list(range(n))
pairs = JA.sweep_pairs(order, inact)[:P]
Ud = dev.to_device(U0)
g = dev.rotate_rdm1(Ud, g, n)
dev.rotate_rdm2_(Ud, G, n)
W = dev.empty(n, n)
dev.sweep_reset(W, n)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
res, summ = dev.sweep_cycle(g, G, Ud.clone(), W, n, dev.to_device(pairs, dev.I32), dev.inactive_mask(inact, n), tables.device_tables(),
                            engine=engine)
e1.record()
torch.cuda.synchronize()
print('accepted', int(summ[0].item()), 'of', len(pairs), 'engine', int(summ[24].item()) or 2, f'{e0.elapsed_time(e1):.2f} ms')
