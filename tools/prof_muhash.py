"""ncu target: one MuHash product tree over N random raw elements (kgv_muhash_elements)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rusty_kaspa_b200 as rk
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rng = np.random.default_rng(1)
items = [rng.integers(0, 256, size=97, dtype=np.uint8).tobytes() for _ in range(n)]
ctx = rk.GpuContext(0)
for _ in range(2):
    m = rk.MuHash(ctx).update(add=items)
print(m.numerator[:8].hex())
