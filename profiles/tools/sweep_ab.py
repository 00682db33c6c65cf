"""A/B timing of the sweep kernel: one full n-orbital cycle (seed-0 order) per library build given on the command line
(each in its own process through QIO_B200_LIB), best and median of `reps` launches.
    python profiles/tools/sweep_ab.py [n] lib1.so lib2.so ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r)
from qio_b200 import device as dev, synth, tables
from qio_b200.grad import jacobi as JA
n = %d
w, Pk = synth.projectors(n)
P = torch.from_numpy(Pk).cuda(); wt = torch.from_numpy(w).cuda()
g1 = torch.einsum('k,kab->ab', wt, P)
gamma0 = torch.zeros((2*n, 2*n), dtype=torch.float64, device='cuda'); gamma0[0::2,0::2] = g1; gamma0[1::2,1::2] = g1
G0 = torch.einsum('k,kad,kbc->abcd', wt, P, P).contiguous()
np.random.seed(0); U0 = JA.random_start(n); order = np.arange(n); np.random.shuffle(order)
inact = list(range(n)); pairs = dev.to_device(JA.sweep_pairs(order, inact), dev.I32)
Ud0 = dev.to_device(U0); mask = dev.inactive_mask(inact, n); tabs = tables.device_tables(); W = dev.empty(n, n)
times = []
for rep in range(%d):
    G = G0.clone(); dev.rotate_rdm2_(Ud0, G, n); g = dev.rotate_rdm1(Ud0, gamma0, n); Ud = Ud0.clone(); dev.sweep_reset(W, n)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    res, summ = dev.sweep_cycle(g, G, Ud, W, n, pairs, mask, tabs)
    b.record(); torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
print('RESULT', min(times), sorted(times)[len(times)//2], int(summ[0].item()))
'''
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100
libs = [a for a in sys.argv[1:] if a.endswith('.so')]
for lib in libs:
    env = dict(os.environ, QIO_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, '-c', CHILD % (ROOT, n, 5)], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith('RESULT')]
    print(os.path.basename(lib), line[0] if line else out.stderr[-400:], flush=True)
