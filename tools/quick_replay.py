"""First timing of kgv_replay_window on a generated chain (Python generator, small): windowed vs blockwise, digest check."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, simgen
from rusty_kaspa_b200.replay import DagReplayer, replay_blocks_array, REPLAY_ACCEPT_COINBASE
from rusty_kaspa_b200.txbatch import build_batch

n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 300
tpb = int(sys.argv[2]) if len(sys.argv) > 2 else 100
win = int(sys.argv[3]) if len(sys.argv) > 3 else 100
t0 = time.perf_counter()
dag = simgen.SimDag(seed=3, n_keys=256, n_nonces=512, coinbase_maturity=5, coinbase_outputs=64, frac_invalid=0.02)
blocks = [dag.make_block(tpb) for _ in range(n_blocks)]
print(f"generated {n_blocks} blocks, {sum(len(b[0]) for b in blocks)} txs in {time.perf_counter() - t0:.1f}s", flush=True)
ctx = rk.GpuContext(0)
prm = Params(coinbase_maturity=5, storage_mass_parameter=dag.C)
r1 = DagReplayer(ctx, prm, 1 << 20)
t0 = time.perf_counter(); got1 = r1.replay_blockwise(blocks); ctx.synchronize(); t_block = time.perf_counter() - t0
r2 = DagReplayer(ctx, prm, 1 << 20)
# pre-flatten windows (host prep is not what is being timed here)
wins = []
for w in range(0, n_blocks, win):
    txs, ranges = [], []
    for t, pov in blocks[w:w + win]:
        ranges.append((len(txs), len(t), pov, REPLAY_ACCEPT_COINBASE)); txs.extend(t)
    wins.append((build_batch(txs), replay_blocks_array(ranges), ranges))
for rep in range(2):
    r2.close(); r2 = DagReplayer(ctx, prm, 1 << 20)
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    got2 = []
    sigs = 0
    for b, arr, ranges in wins:
        res = r2.replay_window(b, arr)
        sigs += r2.last_stats["n_sig_checks"]
        got2 += [res[f:f + n] for f, n, _, _ in ranges]
    t_win = time.perf_counter() - t0
    launches = ctx.launch_count - l0
ntx = sum(len(b[0]) - 1 for b in blocks)
same = all((a["status"] == c["status"]).all() and (a["script_err"] == c["script_err"]).all() for a, c in zip(got1, got2))
print(f"blockwise {t_block*1e3:.1f} ms ({ntx/t_block/1e3:.1f} k tx/s) | windowed({win}) {t_win*1e3:.1f} ms ({ntx/t_win/1e3:.1f} k tx/s, {sigs/t_win/1e6:.2f} M sig/s), {launches} launches, "
      f"verdicts_same={same} digest_same={r1.us.digest() == r2.us.digest()} count={r2.us.count()}", flush=True)
