"""A few prep_rdm12 calls at n orbitals for ncu captures / timing of prep_rdm2_kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qio_b200 import device as dev  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dm1 = torch.randn(n, n, dtype=torch.float64, device='cuda')
dm2 = torch.randn(n, n, n, n, dtype=torch.float64, device='cuda')
Gamma = dev.empty(n, n, n, n)
gamma = dev.empty(2 * n, 2 * n)
dev.prep_rdm12(dm1, dm2, n, gamma=gamma, Gamma=Gamma)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    dev.prep_rdm12(dm1, dm2, n, gamma=gamma, Gamma=Gamma)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'n={n}: prep_rdm12 {ms:.3f} ms = {16 * n ** 4 / ms / 1e6:.0f} GB/s')
