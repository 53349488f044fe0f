"""K8 MuHash on the GPU through the C ABI: the reference's known answers (tests/golden/muhash.json), parity with the
oracle on random element sets and on transaction batches, algebraic properties at larger sizes."""
import ctypes
import json
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

pytestmark = pytest.mark.gpu


class OkMuHash(ctypes.Structure):
    _fields_ = [("num", ctypes.c_uint64 * 48), ("den", ctypes.c_uint64 * 48)]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "muhash.json")))


@pytest.fixture(scope="module")
def ctx():
    import rusty_kaspa_b200 as rk
    c = rk.GpuContext(0)
    yield c
    c.close()


def oracle_raw(oracle, m):
    num, den = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
    oracle.ok_muhash_raw(ctypes.byref(m), num, den)
    return num.raw, den.raw


def test_reference_known_answers(ctx, golden):
    from rusty_kaspa_b200 import MuHash
    assert MuHash(ctx).finalize().hex() == golden["empty_muhash"]                       # test_empty_hash
    cum = MuHash(ctx)
    for v in golden["test_vectors"]:                                                     # test_vectors_hash / add_remove
        d = bytes.fromhex(v["data"])
        assert MuHash(ctx).add_element(d).finalize().hex() == v["multiset_hash"]
        assert cum.add_element(d).finalize().hex() == v["cumulative_hash"]
    for i in reversed(range(3)):
        cum.remove_element(bytes.fromhex(golden["test_vectors"][i]["data"]))
        want = golden["test_vectors"][i - 1]["cumulative_hash"] if i else golden["empty_muhash"]
        assert cum.finalize().hex() == want
    pc = golden["pre_computed"]                                                          # test_new_pre_computed
    m = MuHash(ctx).update(add=[bytes.fromhex(h) for h in pc["add"]], remove=[bytes.fromhex(h) for h in pc["remove"]])
    assert m.finalize().hex() == pc["finalized"]
    se = golden["serialize"]                                                             # test_serialize
    m = MuHash(ctx).update(add=[bytes.fromhex(h) for h in se["add"]])
    assert m.serialize().hex() == se["serialized"]
    # test_vectors_combine_subtract
    m1 = MuHash(ctx).update(add=[bytes.fromhex(v["data"]) for v in golden["test_vectors"]])
    m2 = MuHash(ctx).update(remove=[bytes.fromhex(v["data"]) for v in golden["test_vectors"]])
    m1.combine(m2)
    assert m1.finalize().hex() == golden["empty_muhash"]


def test_random_sets_match_oracle(ctx, oracle):
    from rusty_kaspa_b200 import MuHash
    rnd = random.Random(9)
    for n in (1, 2, 3, 7, 64, 257, 1000):
        items = [bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 200))) for _ in range(n)]
        rem = [rnd.random() < 0.4 for _ in range(n)]
        m = OkMuHash()
        oracle.ok_muhash_init(ctypes.byref(m))
        for d, r in zip(items, rem):
            (oracle.ok_muhash_remove_element if r else oracle.ok_muhash_add_element)(ctypes.byref(m), d, len(d))
        g = MuHash(ctx).update(add=[d for d, r in zip(items, rem) if not r], remove=[d for d, r in zip(items, rem) if r])
        assert (g.numerator, g.denominator) == oracle_raw(oracle, m), n
    out = ctypes.create_string_buffer(32)
    oracle.ok_muhash_finalize(ctypes.byref(m), out)
    assert g.finalize() == out.raw


def test_transaction_batches_match_oracle(ctx, oracle):
    """kgv_muhash_txs (populated entries, and entries looked up in the GPU UTXO table) vs ok_muhash_accepted"""
    from rusty_kaspa_b200 import MuHash, GpuUtxoSet, simgen
    from rusty_kaspa_b200.txbatch import build_batch
    import oracle_tx
    fk, fe, txs = simgen.funded_window(300, n_keys=32, n_nonces=32, mix=(0.5, 0.2, 0.15, 0.15))
    ents, k = [], 0
    for t in txs:
        ents.append(fe[k:k + len(t["inputs"])]); k += len(t["inputs"])
    b = build_batch(txs, ents)
    ob = oracle_tx.ok_batch(b)
    rng = np.random.default_rng(3)
    for frac in (1.0, 0.7, 0.0):
        accept = (rng.random(len(txs)) < frac).astype(np.uint8)
        m = OkMuHash()
        oracle.ok_muhash_accepted(ctypes.byref(m), ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), accept.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(1234))
        g = MuHash.from_transactions(ctx, b, accept, 1234)
        assert (g.numerator, g.denominator) == oracle_raw(oracle, m), frac
    # entries from the table instead of the batch
    us = GpuUtxoSet(ctx, capacity_slots=4096)
    ae, ab = simgen.entries_to_arrays(fe)
    us.apply_diff(add_keys36=fk, add_entries=ae, add_bytes=ab)
    g2 = MuHash.from_transactions(ctx, b, accept * 0 + 1, 1234, utxo_set=us)
    accept1 = np.ones(len(txs), dtype=np.uint8)
    oracle.ok_muhash_accepted(ctypes.byref(m), ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), accept1.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(1234))
    assert (g2.numerator, g2.denominator) == oracle_raw(oracle, m)
    # the reference-shaped entry point: validate + MuHash of the accepted subset (one input signature corrupted -> that tx drops out)
    from rusty_kaspa_b200 import TransactionValidator, Params
    tv = TransactionValidator(ctx, Params(storage_mass_parameter=simgen.DEFAULT_STORAGE_MASS_PARAMETER))
    res, g3 = tv.validate_transactions_with_muhash_in_parallel(us, b, 1234)
    assert (res["status"] == 0).all() and (g3.numerator, g3.denominator) == (g2.numerator, g2.denominator)
    # UTXO-set commitment: set after = set before - spent + created  <=>  H(after) = H(before) * num / den
    before = MuHash.of_utxo_set(ctx, us)
    us.add_transactions(b, accept1, 1234)
    after = MuHash.of_utxo_set(ctx, us)
    before.combine(g2)
    assert before.finalize() == after.finalize()
    us.close()


def test_large_tree_properties(ctx):
    """size-independent properties at a size the oracle would take minutes for: order independence, add-then-remove = empty"""
    from rusty_kaspa_b200 import MuHash
    rng = np.random.default_rng(11)
    n = 40000
    items = [rng.integers(0, 256, size=int(rng.integers(20, 120)), dtype=np.uint8).tobytes() for _ in range(n)]
    a = MuHash(ctx).update(add=items)
    perm = rng.permutation(n)
    b = MuHash(ctx).update(add=[items[i] for i in perm[: n // 2]]).update(add=[items[i] for i in perm[n // 2:]])
    assert a.numerator == b.numerator and a.denominator == b.denominator == (1).to_bytes(384, "little")
    c = MuHash(ctx).update(remove=items)
    a.combine(c)
    assert a.finalize() == MuHash(ctx).finalize()


def test_real_header_commitments_of_the_simpa_dag(ctx):
    """utxoCommitment and acceptedIdMerkleRoot written by the reference into the headers of its simpa DAG fixture, reproduced on the GPU
    along the longest selected-parent chain whose past holds only coinbase transactions: coinbase outputs go into the GPU UTXO table
    (kgv_utxo_apply_accepted with the child's daa score), kgv_utxo_muhash + kgv_muhash_finalize give the commitment, kgv_tx_ids +
    kgv_merkle_roots + one more branch hash the KIP-15 accepted-id root."""
    from rusty_kaspa_b200 import MuHash, GpuUtxoSet
    from rusty_kaspa_b200.txbatch import build_batch
    from golden_util import simpa_dag_coinbase_only_chain_info
    import pyref
    by, eligible, sp = simpa_dag_coinbase_only_chain_info()
    depth = {}
    for h in eligible:
        depth[h] = 0 if sp(h) is None else depth[sp(h)] + 1
    tip = max(eligible, key=lambda h: depth[h])
    chain = [tip]
    while sp(chain[-1]) is not None:
        chain.append(sp(chain[-1]))
    chain.reverse()
    assert len(chain) >= 40
    us = GpuUtxoSet(ctx, capacity_slots=4096)
    running = MuHash(ctx)
    checked = 0
    for parent, child in zip(chain[:-1], chain[1:]):
        cb = build_batch([by[parent]["txs"][0]])
        acc = np.ones(1, dtype=np.uint8)
        # the selected parent's coinbase enters the set from the child's point of view (utxo_validation.rs:117-121)
        running.combine(MuHash.from_transactions(ctx, cb, acc, by[child]["daa_score"], utxo_set=us))
        us.add_transactions(cb, acc, by[child]["daa_score"])
        ids = ctx.tx_ids(cb)
        inner = ctx.merkle_roots(ids, [0, 1])[0].tobytes()
        assert pyref.blake2b_keyed(b"MerkleBranchHash", bytes.fromhex(by[parent]["accepted_id_merkle_root"]) + inner).hex() == by[child]["accepted_id_merkle_root"]
        if checked < 12 or child == tip:  # every finalize is a 3072-bit inversion (~25 ms): sample the chain, always check the tip
            assert MuHash.of_utxo_set(ctx, us).finalize().hex() == by[child]["utxo_commitment"], child
            checked += 1
    assert running.finalize().hex() == by[tip]["utxo_commitment"]
    us.close()
