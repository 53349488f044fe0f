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


@pytest.mark.gpu
def test_block_set_checks(oracle):
    """duplicate-tx / double-spend / chained-tx checks (body_validation_in_isolation.rs:95-131): every block of the simpa DAG
    fixture passes; generated blocks with injected offenders get the oracle's (status, first offender)."""
    import copy
    import ctypes
    import rusty_kaspa_b200 as rk
    from rusty_kaspa_b200 import simgen
    from rusty_kaspa_b200.txbatch import build_batch
    from golden_util import load, tx_from_json
    ctx = rk.GpuContext(0)
    fx = load("simpa_goref_1060.json.gz")
    txs, first = [], [0]
    for b in fx["blocks"]:
        txs += [tx_from_json(t) for t in b["transactions"]]
        first.append(len(txs))
    res = ctx.block_set_checks(build_batch(txs), first)
    assert (res["status"] == 0).all()
    # generated window cut into blocks of 1..40 txs, offenders injected into some blocks
    _, _, wtxs = simgen.funded_window(400, n_keys=16, n_nonces=16)
    rng = np.random.default_rng(4)
    blocks, k = [], 0
    while k < len(wtxs):
        n = int(rng.integers(1, 40))
        blocks.append([copy.deepcopy(t) for t in wtxs[k:k + n]])
        k += n
    for bi, blk in enumerate(blocks):
        kind = bi % 5
        if kind == 1 and len(blk) >= 2:      # duplicate transaction
            blk.append(copy.deepcopy(blk[int(rng.integers(0, len(blk)))]))
        elif kind == 2 and len(blk) >= 2:    # double spend: a later tx re-spends an earlier outpoint
            a, c = sorted(rng.choice(len(blk), size=2, replace=False))
            blk[c]["inputs"][0]["txid"], blk[c]["inputs"][0]["index"] = blk[a]["inputs"][-1]["txid"], blk[a]["inputs"][-1]["index"]
        elif kind == 3 and len(blk) >= 2:    # chained: spends an output created in the same block (any position)
            a, c = rng.choice(len(blk), size=2, replace=False)
            blk[c]["inputs"][-1]["txid"], blk[c]["inputs"][-1]["index"] = simgen.tx_id(blk[a]), int(rng.integers(0, len(blk[a]["outputs"])))
        elif kind == 4 and len(blk) >= 3:    # both a double spend and a chained tx: the double spend is reported
            blk[2]["inputs"][0]["txid"], blk[2]["inputs"][0]["index"] = blk[0]["inputs"][0]["txid"], blk[0]["inputs"][0]["index"]
            blk[1]["inputs"][0]["txid"], blk[1]["inputs"][0]["index"] = simgen.tx_id(blk[0]), 0
    flat, first = [], [0]
    for blk in blocks:
        flat += blk
        first.append(len(flat))
    batch = build_batch(flat)
    got = ctx.block_set_checks(batch, first)
    ob = oracle_tx.ok_batch(batch)
    seen = set()
    for bi in range(len(blocks)):
        idx = ctypes.c_uint32()
        st = oracle.ok_block_set_checks(ctypes.byref(ob), ctypes.c_uint32(first[bi]), ctypes.c_uint32(first[bi + 1]), ctypes.byref(idx))
        assert (int(got[bi]["status"]), int(got[bi]["index"])) == (st, idx.value if st else 0), bi
        seen.add(st)
    assert seen == {0, 1, 2, 3}
    ctx.close()


def test_body_validation_example_block_of_the_reference_on_the_gpu(gpu_ctx):
    """the reference's own test block (body_validation_in_isolation.rs:153-462): kgv_block_hash_merkle_roots reproduces the header's hash_merkle_root,
    kgv_block_set_checks passes it and reports the test's three mutations as DuplicateTransactions / DoubleSpendInSameBlock / ChainedTransaction -
    all four blocks in one call"""
    from golden_util import body_validation_blocks
    from rusty_kaspa_b200.txbatch import build_batch
    root, blocks = body_validation_blocks()
    flat, first = [], [0]
    for _, txs, _ in blocks:
        flat += txs
        first.append(len(flat))
    b = build_batch(flat)
    assert gpu_ctx.block_hash_merkle_roots(b, first)[0].tobytes().hex() == root
    got = gpu_ctx.block_set_checks(b, first)
    assert got["status"].tolist() == [w for _, _, w in blocks]
    assert int(got[1]["index"]) == first[1] + 5 and int(got[3]["index"]) > 0  # the pushed clone is transaction 5 of its block (absolute index)
