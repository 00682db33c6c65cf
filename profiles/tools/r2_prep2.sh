#!/bin/bash
bash profiles/tools/r2_prep.sh p5
echo "== TMA kernel forced for all n (lib_prepbig)"
QIO_B200_LIB=$PWD/profiles/tools/_exp/lib_prepbig.so python - <<PY
import torch
from qio_b200 import device as dev
for n in (100, 200):
    dm2 = torch.randn((n,n,n,n), dtype=torch.float64, device='cuda'); G = dev.empty(n,n,n,n)
    for _ in range(3): dev.prep_rdm12(None, dm2, n, n, Gamma=G)
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dev.prep_rdm12(None, dm2, n, n, Gamma=G)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)/10
    print(f'n={n}: prep {ms:.4f} ms = {16*n**4/ms/1e9:.2f} TB/s ({16*n**4/ms/1e9/6.4512:.3f})')
    del dm2, G
PY
