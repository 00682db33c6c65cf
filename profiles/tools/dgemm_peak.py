"""cuBLAS DGEMM throughput (the denominator for the FP64 tensor-pipe roofline of rotate_axis)."""
import torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device='cuda'); b = torch.randn(n, n, dtype=torch.float64, device='cuda')
for _ in range(2): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f'cuBLAS DGEMM {n}^3: {best:.2f} ms  {2*n**3/best/1e9:.2f} TFLOP/s')
# tall-skinny like the rotation pass: (n^3 x n) @ (n x n), n = 100
m = 100
x = torch.randn(m**3, m, dtype=torch.float64, device='cuda'); u = torch.randn(m, m, dtype=torch.float64, device='cuda')
for _ in range(2): (x @ u)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = x @ u; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f'cuBLAS DGEMM ({m}^3 x {m}) @ ({m} x {m}): {best:.3f} ms  {2*m**5/best/1e9:.2f} TFLOP/s')
