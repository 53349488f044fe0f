"""Where the end-to-end time of kgv_replay_window goes: wall clock of the prefetch and replay calls vs the device times the call reports
(kgv_replay_stats), page-locked host arrays, with and without kgv_batch_prefetch."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, simgen, GpuUtxoSet
from rusty_kaspa_b200.replay import REPLAY_BLOCK_DTYPE, ReplayStats
from rusty_kaspa_b200.validator import RESULT_DTYPE
from rusty_kaspa_b200.verifier import _KgvTxBatch
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 5
g = simgen.FastDag(seed=1, n_keys=1024, n_nonces=4096, frac_invalid=0.01, coinbase_outputs=16)
ctx = rk.GpuContext(0)
prm = Params(coinbase_maturity=g.maturity, storage_mass_parameter=g.C)
cudart = torch.cuda.cudart()
wins = []
for w in range(n_win + 1):
    g.generate(400 if w == 0 else 1024, 150)
    b, first, pov = g.take()
    arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE)
    arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
    res = np.zeros(len(b.txs), dtype=RESULT_DTYPE)
    for a in (b.txs, b.inputs, b.outputs, b.arena, res):
        assert int(cudart.cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0
    cb = _KgvTxBatch(b.txs.ctypes.data, len(b.txs), b.inputs.ctypes.data, len(b.inputs), b.outputs.ctypes.data, len(b.outputs), None, b.arena.ctypes.data, len(b.arena))
    wins.append((b, arr, res, cb))
lib, h = ctx._lib, ctx._h
for mode in ("no prefetch", "prefetch", "prefetch"):
    us = GpuUtxoSet(ctx, 1 << 24)
    st = ReplayStats()
    rows = []
    for wi, (b, arr, res, cb) in enumerate(wins):
        t0 = time.perf_counter()
        if mode == "prefetch" and wi + 1 < len(wins):
            ctx._check(lib.kgv_batch_prefetch(h, C.byref(wins[wi + 1][3])))
        t1 = time.perf_counter()
        ctx._check(lib.kgv_replay_window(h, us._h, C.byref(cb), arr.ctypes.data, len(arr), C.byref(prm), res.ctypes.data, None, C.byref(st)))
        t2 = time.perf_counter()
        rows.append((len(b.txs), (t1 - t0) * 1e3, (t2 - t1) * 1e3, float(st.pre_check_ms), float(st.in_order_ms)))
    us.close()
    print(mode)
    for r in rows[1:]:
        print("  txs %d: prefetch call %.2f ms, replay call %.2f ms wall (device: pre-check %.2f + in-order %.2f = %.2f ms) -> %.2f ms outside the device phases"
              % (r[0], r[1], r[2], r[3], r[4], r[3] + r[4], r[2] - r[3] - r[4]))
