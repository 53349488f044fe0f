import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import workload as W
H = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "hostsim", "libhostsim_secp.so"))
pk, msg, sig, kind = W.schnorr_triples(4, seed=1, n_keys=4, n_nonces=4, frac_bitflip=0, frac_adversarial=0)
ctx = rk.GpuContext(0)
for i in range(2):
    st, tr = ctx.debug_schnorr_trace(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes())
    ht = np.zeros((32, 16), dtype=np.uint32)
    hst = H.hs_schnorr_trace(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes(), ht.ctypes.data_as(ctypes.c_void_p))
    print("item", i, "device status", st, "host status", hst)
    for s in range(32):
        if not (tr[s] == ht[s]).all():
            print(" stage", s, "DIFF\n  dev ", " ".join("%08x" % x for x in tr[s]), "\n  host", " ".join("%08x" % x for x in ht[s]))
        elif ht[s].any():
            print(" stage", s, "same")
