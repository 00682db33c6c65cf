"""A few rotate_rdm2 calls (4 single-axis DMMA passes each) for ncu captures / timing of the rotation kernels."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qio_b200 import device as dev  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rng = np.random.default_rng(0)
Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
U = dev.to_device(Q)
G = torch.randn(n, n, n, n, dtype=torch.float64, device='cuda')
work = torch.empty_like(G)
dev.rotate_rdm2_(U, G, n, work)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dev.rotate_rdm2_(U, G, n, work)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'n={n}: rotate_rdm2 {ms:.3f} ms = {8 * n ** 5 / ms / 1e9:.2f} TFLOP/s (4 passes, {ms / 4:.3f} ms each)')
