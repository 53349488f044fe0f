import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, TransactionValidator, GpuUtxoSet
from rusty_kaspa_b200.txbatch import build_batch
from golden_util import load, tx_from_json, entry_from_json
mode = sys.argv[1]
ctx = rk.GpuContext(0)
c = load("check_scripts_kat.json")["cases"][0]
tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
b = build_batch([tx], [entries])
if mode == "direct":
    msg = ctx.sighash(b, [(0, 0, 1, False)])
    ss = tx["inputs"][0]["sigscript"]; spk = entries[0]["script"]
    pk = np.frombuffer(spk[1:33], dtype=np.uint8).reshape(1, 32).copy()
    sig = np.frombuffer(ss[1:65], dtype=np.uint8).reshape(1, 64).copy()
    for n in (1, 3):
        st = ctx.verify_schnorr_batch(np.tile(pk, (n, 1)), np.tile(msg, (n, 1)), np.tile(sig, (n, 1)))
        print("direct verify n=%d" % n, st)
elif mode == "fused1":
    tv = TransactionValidator(ctx, Params(coinbase_maturity=100, storage_mass_parameter=0))
    print("fused single tx:", tv.validate_populated_transactions(b, 10**9, flags=2))
elif mode == "fusedsim":
    from rusty_kaspa_b200.simgen import SimDag
    dag = SimDag(seed=11, n_keys=16, n_nonces=16, coinbase_maturity=0, coinbase_outputs=4)
    tv = TransactionValidator(ctx, Params(coinbase_maturity=0, storage_mass_parameter=dag.C))
    us = GpuUtxoSet(ctx, 1 << 10)
    for _ in range(3):
        txs, pov = dag.make_block(4)
        bb = build_batch(txs)
        res = tv.validate_transactions_in_parallel(us, bb, pov)
        print("block", pov, res["status"], res["script_err"])
        us.add_transactions(bb, np.array([1 if (i == 0 or res[i]["status"] == 0) else 0 for i in range(len(txs))], dtype=np.uint8), pov)
        print("count", us.count())
