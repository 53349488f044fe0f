"""Per-phase timing of kgv_replay_window (kgv_replay_stats.pre_check_ms / in_order_ms) on a generated chain."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, simgen
from rusty_kaspa_b200.replay import DagReplayer, REPLAY_BLOCK_DTYPE
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 4
win = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = simgen.FastDag(seed=1, n_keys=1024, n_nonces=4096, frac_invalid=0.01, coinbase_outputs=16, coinbase_maturity=50)
g.generate(400, 150); g.take()  # ramp
ctx = rk.GpuContext(0)
prm = Params(coinbase_maturity=50, storage_mass_parameter=g.C)
# the ramp is replayed too so that the table holds the spendable set
g2 = simgen.FastDag(seed=1, n_keys=1024, n_nonces=4096, frac_invalid=0.01, coinbase_outputs=16, coinbase_maturity=50)
r = DagReplayer(ctx, prm, 1 << 24)
g2.generate(400, 150); b, first, pov = g2.take()
arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE); arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
r.replay_window(b, arr)
for w in range(n_win):
    g2.generate(win, 150); b, first, pov = g2.take()
    arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE); arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
    t0 = time.perf_counter(); r.replay_window(b, arr); dt = time.perf_counter() - t0
    s = r.last_stats
    print(f"window {w}: {len(pov)} blocks {len(b.txs)} txs {s['n_sig_checks']} sig checks: wall {dt*1e3:.2f} ms, pre-check {s['pre_check_ms']:.2f} ms, in-order {s['in_order_ms']:.2f} ms "
          f"({s['in_order_ms']*1e3/len(pov):.1f} us/block)", flush=True)
