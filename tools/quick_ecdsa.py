"""Scratch timing of the ECDSA kernel with device-resident inputs (not the contract bench)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import workload as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pk, msg, sig, kind = W.ecdsa_triples(min(n, 1 << 15), seed=5, n_keys=4096, n_nonces=4096)
reps_t = (n + len(pk) - 1) // len(pk)
pk, msg, sig, kind = (np.tile(a, (reps_t,) + (1,) * (a.ndim - 1))[:n] for a in (pk, msg, sig, kind))
ctx = rk.GpuContext(0)
ctx.use_torch_stream()
dpk, dmsg, dsig = (torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (pk, msg, sig))
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    ctx.verify_ecdsa_batch(dpk, dmsg, dsig, n=n, status=dst)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
ev0.record()
for _ in range(reps):
    ctx.verify_ecdsa_batch(dpk, dmsg, dsig, n=n, status=dst)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
st = dst.cpu().numpy()
print(f"ecdsa n={n} {ms:.2f} ms/batch  {n / ms * 1e3 / 1e6:.2f} M verifies/s  valid={int((st == 1).sum())} expected_valid={int((kind == 0).sum())}")
