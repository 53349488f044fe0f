"""GPU parity: kgv_tx_ids / kgv_tx_hashes / kgv_sighash (C ABI) vs the reference's vectors and the C oracle."""
import numpy as np
import pytest

import oracle_tx
from golden_util import apply_sighash_action, entry_from_json, load, tx_from_json
from rusty_kaspa_b200 import workload as W
from rusty_kaspa_b200.txbatch import build_batch

pytestmark = pytest.mark.gpu
HASH_TYPES = [1, 2, 4, 0x81, 0x82, 0x84]


def test_tx_id_and_hash_reference_vectors(gpu_ctx):
    vec = load("tx_hashing.json")["vectors"]
    b = build_batch([tx_from_json(v["tx"]) for v in vec])
    ids, hashes = gpu_ctx.tx_ids(b), gpu_ctx.tx_hashes(b)
    for i, v in enumerate(vec):
        assert ids[i].tobytes().hex() == v["expected_id"]
        assert hashes[i].tobytes().hex() == v["expected_hash"]


def test_sighash_reference_vectors(gpu_ctx):
    g = load("sighash.json")
    for v in g["vectors"]:
        tx = tx_from_json(g[v["tx"]])
        entries = [entry_from_json(e) for e in g["entries"]]
        apply_sighash_action(tx, entries, v["action"], v["action_arg"])
        b = build_batch([tx], [entries])
        out = gpu_ctx.sighash(b, [(0, v["input_index"], v["hash_type"], False)])
        assert out[0].tobytes().hex() == v["expected"], v["name"]


def test_hashing_parity_random_transactions(gpu_ctx, oracle):
    txs, entries = W.random_transactions(3000, seed=9)
    b = build_batch(txs, entries)
    assert (gpu_ctx.tx_ids(b) == oracle_tx.tx_ids(oracle, b, threads=8)).all()
    assert (gpu_ctx.tx_hashes(b) == oracle_tx.tx_hashes(oracle, b, threads=8)).all()
    rng = np.random.default_rng(1)
    items = []
    for ti, t in enumerate(txs):
        first = int(b.txs[ti]["first_input"])
        for k in range(len(t["inputs"])):
            items.append((ti, first + k, HASH_TYPES[int(rng.integers(0, 6))], bool(rng.integers(0, 2))))
    got = gpu_ctx.sighash(b, items)
    for i, (ti, a, h, e) in enumerate(items):
        rel = a - int(b.txs[ti]["first_input"])
        assert got[i].tobytes() == oracle_tx.sighash(oracle, b, ti, rel, h, ecdsa=e), (ti, rel, h, e)


def test_sighash_rejects_unknown_hash_type(gpu_ctx):
    txs, entries = W.random_transactions(4, seed=3)
    txs = [t for t in txs if t["inputs"]] or txs
    b = build_batch(txs, [e for t, e in zip(*W.random_transactions(4, seed=3)) if t["inputs"]] or entries)
    out = gpu_ctx.sighash(b, [(0, int(b.txs[0]["first_input"]), 3, False)])
    assert (out[0] == 0xFF).all()


def test_empty_batch(gpu_ctx):
    b = build_batch([])
    assert gpu_ctx.tx_ids(b).shape == (0, 32)


@pytest.mark.gpu
def test_merkle_roots_match_the_fixture_and_the_oracle(oracle):
    """kgv_block_hash_merkle_roots on the 266 blocks of the simpa DAG fixture (header hashMerkleRoot), kgv_merkle_roots on ragged
    random groups (empty, single, odd, powers of two +-1, 1000) against the oracle."""
    import ctypes
    import random
    import rusty_kaspa_b200 as rk
    from rusty_kaspa_b200.txbatch import build_batch
    from golden_util import load, tx_from_json
    fx = load("simpa_goref_1060.json.gz")
    txs, first = [], [0]
    for b in fx["blocks"]:
        txs += [tx_from_json(t) for t in b["transactions"]]
        first.append(len(txs))
    ctx = rk.GpuContext(0)
    roots = ctx.block_hash_merkle_roots(build_batch(txs), first)
    assert [r.tobytes().hex() for r in roots] == [b["hash_merkle_root"] for b in fx["blocks"]]
    rnd = random.Random(8)
    sizes = [0, 1, 2, 3, 4, 5, 7, 8, 9, 0, 31, 32, 33, 1000, 1, 64, 65, 127]
    hs = np.frombuffer(bytes(rnd.randrange(256) for _ in range(32 * sum(sizes))), dtype=np.uint8).reshape(-1, 32)
    f = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint32)
    got = ctx.merkle_roots(hs, f)
    for g, n in enumerate(sizes):
        out = ctypes.create_string_buffer(32)
        oracle.ok_merkle_root(hs[f[g]:f[g + 1]].tobytes(), ctypes.c_size_t(n), out)
        assert got[g].tobytes() == out.raw, (g, n)
    ctx.close()
