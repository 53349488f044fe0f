"""Times k_schnorr_verify for several builds of libkgv (variants/libkgv_*.so via KGV_LIB), device-resident inputs."""
import sys, os, glob, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import rusty_kaspa_b200 as rk
    from rusty_kaspa_b200 import workload as W
    n = 1 << 20
    pk, msg, sig, kind = W.schnorr_triples(1 << 15, seed=5, n_keys=4096, n_nonces=4096)
    pk, msg, sig, kind = W.tile_triples(pk, msg, sig, kind, n)
    ctx = rk.GpuContext(0)
    s = torch.cuda.Stream()
    ctx.use_stream(s.cuda_stream)
    with torch.cuda.stream(s):
        dpk, dmsg, dsig = (torch.from_numpy(a).cuda() for a in (pk, msg, sig))
        dst = torch.empty(n, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            ctx.verify_schnorr_batch(dpk, dmsg, dsig, n=n, status=dst)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(3):
            ctx.verify_schnorr_batch(dpk, dmsg, dsig, n=n, status=dst)
        e1.record(s)
        s.synchronize()
    ms = e0.elapsed_time(e1) / 3
    st = dst.cpu().numpy()
    ok = int((st == 1).sum()) == int((kind == 0).sum()) and not (st[kind != 0] == 1).any()
    print(f"{os.path.basename(os.environ.get('KGV_LIB', 'default'))}: {ms:.2f} ms  {n / ms / 1e3:.2f} M/s  verdicts_ok={ok}", flush=True)
else:
    for v in sorted(glob.glob(os.path.join(ROOT, "variants", "libkgv_*.so"))):
        subprocess.run(["timeout", "120", sys.executable, __file__, "--one"], env=dict(os.environ, KGV_LIB=v))
