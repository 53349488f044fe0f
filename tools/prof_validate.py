"""ncu target: one kgv_validate_txs + kgv_muhash_txs + kgv_utxo_apply_accepted on a window of N independent txs (default 32768)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import simgen
from rusty_kaspa_b200.txbatch import build_batch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
mix = (1.0, 0, 0, 0) if len(sys.argv) < 3 else (0.5, 0.0, 0.25, 0.25)
fk, fe, txs = simgen.funded_window(n, mix=mix)
b = build_batch(txs)
ae, ab = simgen.entries_to_arrays(fe)
ctx = rk.GpuContext(0)
us = rk.GpuUtxoSet(ctx, 4 * len(fk))
us.apply_diff(add_keys36=fk, add_entries=ae, add_bytes=ab)
tv = rk.TransactionValidator(ctx, rk.Params(coinbase_maturity=100, storage_mass_parameter=simgen.DEFAULT_STORAGE_MASS_PARAMETER))
for _ in range(2):
    t0 = time.perf_counter()
    res, mh = tv.validate_transactions_with_muhash_in_parallel(us, b, 10)
    print("validate+muhash wall ms", (time.perf_counter() - t0) * 1e3, int((res["status"] == 0).sum()))
us.add_transactions(b, (res["status"] == 0).astype(np.uint8), 10)
print(us.count())
