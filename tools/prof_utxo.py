"""Short single-GPU run for ncu: K5 on its own - insert 2 Mi entries into an 8 Mi-slot table, look every entry up in a random order (twice),
erase and re-insert 512 Ki of them.  Kernels: k_utxo_insert, k_utxo_lookup, k_utxo_erase."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import GpuUtxoSet
from rusty_kaspa_b200.txbatch import ENTRY_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
rng = np.random.default_rng(7)
keys = rng.integers(0, 256, size=(n, 36), dtype=np.uint8)
ent = np.zeros(n, dtype=ENTRY_DTYPE)
ent["amount"] = rng.integers(1, 1 << 40, size=n)
ent["script_off"] = (np.arange(n, dtype=np.uint64) * 34 % (1 << 20)).astype(np.uint32)
ent["script_len"] = 34
arena = rng.integers(0, 256, size=(1 << 20) + 64, dtype=np.uint8)
ctx = rk.GpuContext(0)
us = GpuUtxoSet(ctx, 4 * n)
us.apply_diff(add_keys36=keys, add_entries=ent, add_bytes=arena)
dev = torch.device("cuda:0")
dkeys = torch.from_numpy(keys).to(dev)
de = torch.empty(n * ENTRY_DTYPE.itemsize, dtype=torch.uint8, device=dev)
df = torch.empty(n, dtype=torch.uint8, device=dev)
for r in range(2):
    dk = dkeys[torch.randperm(n, device=dev)].contiguous()
    ctx._check(ctx._lib.kgv_utxo_lookup(ctx._h, us._h, dk.data_ptr(), n, de.data_ptr(), None, 0, df.data_ptr()))
ctx.synchronize()
assert int(df.sum().item()) == n
m = n // 4
us.apply_diff(rem_keys36=keys[:m])
us.apply_diff(add_keys36=keys[:m], add_entries=ent[:m], add_bytes=arena)
print("entries", us.count())
