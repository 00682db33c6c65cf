"""Arrival / release times of one mid-sweep grid barrier per CTA (debug aid for the persistent sweep kernel)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from qio_b200 import device as dev, synth, tables
from qio_b200._lib import lib, LS_DTYPE, check
from qio_b200.grad import jacobi as JA
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
gamma, Gamma = synth.prepared_rdms(n)
np.random.seed(0); U0 = JA.random_start(n); order = np.arange(n); np.random.shuffle(order)
inact = list(range(n)); pairs = JA.sweep_pairs(order, inact)
g, G, Ud = dev.to_device(gamma), dev.to_device(Gamma), dev.to_device(U0)
g = dev.rotate_rdm1(Ud, g, n); dev.rotate_rdm2_(Ud, G, n)
W = dev.empty(n, n); dev.sweep_reset(W, n)
P = len(pairs); pd = dev.to_device(pairs, dev.I32); mask = dev.inactive_mask(inact, n); tabs = tables.device_tables()
res = torch.empty(P * LS_DTYPE.itemsize, dtype=torch.uint8, device='cuda'); summ = dev.empty(32)
wsb = int(lib.qio_b200_sweep_workspace_bytes(n)); ws = torch.zeros(wsb, dtype=torch.uint8, device='cuda')
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
check(lib.qio_b200_sweep_cycle(g.data_ptr(), G.data_ptr(), Ud.data_ptr(), W.data_ptr(), n, 0, n, pd.data_ptr(), P, mask.data_ptr(),
                               ctypes.byref(tabs.struct), res.data_ptr(), summ.data_ptr(), ws.data_ptr(), None, st))
torch.cuda.synchronize()
dbg = ws[wsb - 2 * 1024 * 8:].view(torch.float64).cpu().numpy()
sms = torch.cuda.get_device_properties(0).multi_processor_count
arr, rel = dbg[:sms], dbg[1024:1024 + sms]
t0 = arr.min()
print('arrival us (sorted, last 12):', np.round((np.sort(arr) - t0)[-12:] / 1e3, 2))
print('latest arrivers (cta ids):', np.argsort(arr)[-8:])
print('release us: min %.2f max %.2f ; last arrival %.2f' % ((rel.min() - t0) / 1e3, (rel.max() - t0) / 1e3, (arr.max() - t0) / 1e3))
print('arrival of cta0 %.2f, cta1 %.2f, cta %d (flat) %.2f, gamma cta %.2f, wu cta %.2f' % ((arr[0]-t0)/1e3, (arr[1]-t0)/1e3, sms-3, (arr[sms-3]-t0)/1e3, (arr[sms-1]-t0)/1e3, (arr[sms-2]-t0)/1e3))
