#!/bin/bash
# prep kernel timing at n = 100 / 200 (+ the parity tests of prep)
tag=${1:-p}; out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prep or div6" > $out/${tag}_pytest.log 2>&1; tail -2 $out/${tag}_pytest.log
python - <<PY
import torch, time
from qio_b200 import device as dev
for n in (100, 200):
    dm2 = torch.randn((n,n,n,n), dtype=torch.float64, device='cuda')
    G = dev.empty(n,n,n,n)
    for _ in range(3): dev.prep_rdm12(None, dm2, n, n, Gamma=G)
    torch.cuda.synchronize()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): dev.prep_rdm12(None, dm2, n, n, Gamma=G)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)/10
    print(f'n={n}: prep {ms:.4f} ms = {16*n**4/ms/1e9:.2f} TB/s ({16*n**4/ms/1e9/6.4512:.3f} of 6451 GB/s)')
    del dm2, G
PY
