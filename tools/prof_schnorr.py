"""Short single-GPU run for ncu: a few launches of k_schnorr_verify at a given batch size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import workload as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
pk, msg, sig, kind = W.schnorr_triples(min(n, 1 << 16), seed=5, n_keys=4096, n_nonces=4096)
pk, msg, sig, kind = W.tile_triples(pk, msg, sig, kind, n)
ctx = rk.GpuContext(0)
dpk, dmsg, dsig = (torch.from_numpy(a).cuda() for a in (pk, msg, sig))
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for _ in range(reps):
    ctx.verify_schnorr_batch(dpk, dmsg, dsig, n=n, status=dst)
ctx.synchronize()
print("valid", int((dst.cpu().numpy() == 1).sum()), "of", n)
