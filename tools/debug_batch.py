import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import workload as W
pk, msg, sig, kind = W.schnorr_triples(300, seed=1, n_keys=4, n_nonces=4, frac_bitflip=0, frac_adversarial=0)
ctx = rk.GpuContext(0)
for n in [1, 2, 32, 33, 128, 129, 300]:
    st = ctx.verify_schnorr_batch(pk[:n].copy(), msg[:n].copy(), sig[:n].copy(), n=n)
    print("host path n=%d" % n, np.bincount(st, minlength=4))
ctx.use_torch_stream()
for off in [0, 1]:
    n = 300
    bufs = []
    for a in (pk, msg, sig):
        t = torch.zeros(a.size + 64, dtype=torch.uint8, device="cuda")
        t[off:off + a.size] = torch.from_numpy(a.reshape(-1)).cuda()
        bufs.append(t[off:off + a.size])
    st = torch.full((n,), 9, dtype=torch.uint8, device="cuda")
    ctx.verify_schnorr_batch(bufs[0], bufs[1], bufs[2], n=n, status=st)
    torch.cuda.synchronize()
    print("device path offset=%d" % off, np.bincount(st.cpu().numpy(), minlength=10))
