"""Multi-GPU exchange through the C ABI (kgv_comm_*, kgv_shard_*, kgv_set_sharding).  On a one-GPU box the peer transport is
exercised with several contexts of one process on the same device (kgv_comm_connect_local; one host thread per rank, as a host
would drive several GPUs); with two or more GPUs the same tests place one context per device.  The NCCL transport needs one
process per GPU: tools/mgpu_check.py (run under torchrun) covers it."""
import threading

import numpy as np
import pytest
import torch

import oracle_tx
import rusty_kaspa_b200 as rk
from rusty_kaspa_b200 import Params, workload as W
from rusty_kaspa_b200.comm import ShardComm

pytestmark = pytest.mark.gpu


def _contexts(n):
    ndev = torch.cuda.device_count()
    return [rk.GpuContext(r % ndev) for r in range(n)]


def _run_ranks(fn, n):
    errs, out = [], [None] * n

    def body(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # noqa: BLE001
            errs.append((r, e))
    th = [threading.Thread(target=body, args=(r,)) for r in range(n)]
    [t.start() for t in th]
    [t.join(timeout=120) for t in th]
    assert not errs, errs
    assert not any(t.is_alive() for t in th), "a rank did not finish (deadlock?)"
    return out


@pytest.mark.parametrize("n_ranks", [2, 4])
def test_peer_exchange_of_shard_bitmaps(n_ranks):
    """every rank verifies its own shard, kgv_shard_publish_bitmap writes the packed verdicts into every peer, kgv_shard_wait collects:
    all ranks hold the global bitmap = concatenation of the per-shard bitmaps the single-GPU path produces; repeated for several epochs"""
    n = 4096
    ctxs = _contexts(n_ranks)
    shards = [W.schnorr_triples(n, seed=100 + r, n_keys=64, n_nonces=64, frac_bitflip=0.1, frac_adversarial=0.1) for r in range(n_ranks)]
    comms = [ShardComm(ctxs[r], n_ranks, r, slice_capacity=n // 8) for r in range(n_ranks)]
    ShardComm.connect_local(comms)
    expect = np.concatenate([np.packbits((ctxs[0].verify_schnorr_batch(*s[:3]) == 1).astype(np.uint8), bitorder="little") for s in shards])

    def rank_body(r):
        ctx, c = ctxs[r], comms[r]
        dev = torch.device("cuda", ctx.device)
        pk, msg, sig, _ = shards[r]
        d = [torch.from_numpy(a).to(dev) for a in (pk, msg, sig)]
        st = torch.empty(n, dtype=torch.uint8, device=dev)
        allb = torch.zeros(n_ranks * n // 8, dtype=torch.uint8, device=dev)
        got = []
        for _ in range(5):
            ctx.verify_schnorr_batch(d[0], d[1], d[2], n=n, status=st)
            e = c.publish_bitmap(st.data_ptr(), n)
            c.wait(e, n // 8, allb.data_ptr())
            ctx.synchronize()
            got.append(allb.cpu().numpy().copy())
        return got
    res = _run_ranks(rank_body, n_ranks)
    for r in range(n_ranks):
        for g in res[r]:
            assert (g == expect).all(), r
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()


def test_sharded_replay_equals_unsharded(oracle):
    """kgv_set_sharding: two ranks replay the same window against their own table replica, each verifying half of the candidate
    (signature, key) pairs and exchanging the verdicts; results, accepted sets and UTXO digests equal the oracle's (and so the unsharded run's)"""
    from rusty_kaspa_b200 import simgen
    from rusty_kaspa_b200.replay import DagReplayer, REPLAY_BLOCK_DTYPE
    n_ranks = 2
    g = simgen.FastDag(seed=21, n_keys=64, n_nonces=256, coinbase_maturity=3, mix=(0.4, 0.2, 0.2, 0.2), frac_invalid=0.1, coinbase_outputs=12)
    wins = []
    for _ in range(3):
        g.generate(40, 40)
        wins.append(g.take())
    prm = Params(coinbase_maturity=3, storage_mass_parameter=g.C)
    ost = oracle_tx.State(oracle)
    op = oracle_tx.params(coinbase_maturity=3, storage_mass_parameter=g.C)
    exp = [oracle_tx.state_replay(ost, b, first, pov, op, threads=8) for b, first, pov in wins]
    ctxs = _contexts(n_ranks)
    comms = [ShardComm(ctxs[r], n_ranks, r, slice_capacity=1 << 20) for r in range(n_ranks)]
    ShardComm.connect_local(comms)
    reps = [DagReplayer(ctxs[r], prm, 1 << 16) for r in range(n_ranks)]
    # On a one-GPU box the "ranks" are contexts of ONE device: size every per-call buffer up front (an unsharded dry run on a scratch table), because
    # growing one later means cudaMalloc while the other rank's wait kernel spins - harmless across devices / processes, the real deployment.
    if torch.cuda.device_count() < n_ranks:
        for r in range(n_ranks):
            scratch = DagReplayer(ctxs[r], prm, 1 << 16)
            for b, first, pov in wins:  # (buffer sizes follow txs / inputs / outputs / signatures of a window, rounded to powers of two in places)
                arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE)
                arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
                scratch.replay_window(b, arr)
            scratch.close()

    def rank_body(r):
        comms[r].shard_validation(True)
        out = []
        for b, first, pov in wins:
            arr = np.zeros(len(pov), dtype=REPLAY_BLOCK_DTYPE)
            arr["first_tx"], arr["n_txs"], arr["pov_daa_score"], arr["flags"] = first[:-1], np.diff(first), pov, 1
            out.append(reps[r].replay_window(b, arr, want_accept=True))
        return out, reps[r].us.count(), reps[r].us.digest()
    res = _run_ranks(rank_body, n_ranks)
    for r in range(n_ranks):
        out, cnt, dig = res[r]
        for (got, acc), (e, eacc) in zip(out, exp):
            assert (got["status"] == e["status"]).all() and (got["script_err"] == e["script_err"]).all() and (acc == eacc).all()
        assert cnt == ost.count() and dig == ost.digest()
    for c in comms:
        c.close()
    for rp in reps:
        rp.close()
    for c in ctxs:
        c.close()
    ost.close(); g.close()
