import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np, pyref
import rusty_kaspa_b200 as rk
ctx = rk.GpuContext(0)
P = pyref.MUHASH_P
m = rk.MuHash(ctx)
print("empty finalize", m.finalize().hex(), pyref.MuHash().finalize().hex())
import random
rnd = random.Random(1)
for it in range(4):
    a, b, c, d = [rnd.getrandbits(3072) % P for _ in range(4)]
    x = rk.MuHash(ctx, a.to_bytes(384, "little"), b.to_bytes(384, "little"))
    y = rk.MuHash(ctx, c.to_bytes(384, "little"), d.to_bytes(384, "little"))
    x.combine(y)
    print("combine ok:", int.from_bytes(x.numerator, "little") == a * c % P, int.from_bytes(x.denominator, "little") == b * d % P)
    ser = x.serialize()
    want = (a * c % P) * pow(b * d % P, P - 2, P) % P
    print("finalize ok:", int.from_bytes(ser, "little") == want)
