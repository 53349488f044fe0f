"""MuHash: the C oracle (oracle/ok_muhash.c) and its Python twin against the reference's own known answers
(tests/golden/muhash.json, extracted from crypto/muhash/src/lib.rs by tests/golden/make_golden.py), then against each
other on random data, then the transaction-level restatement (consensus/core/src/muhash.rs)."""
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
import pyref  # noqa: E402


class OkMuHash(ctypes.Structure):
    _fields_ = [("num", ctypes.c_uint64 * 48), ("den", ctypes.c_uint64 * 48)]


@pytest.fixture(scope="module")
def golden():
    return json.load(open(os.path.join(HERE, "golden", "muhash.json")))


def c_finalize(lib, m):
    out = ctypes.create_string_buffer(32)
    lib.ok_muhash_finalize(ctypes.byref(m), out)
    return out.raw


def new(lib):
    m = OkMuHash()
    lib.ok_muhash_init(ctypes.byref(m))
    return m


def test_reference_known_answers(oracle, golden):
    assert golden["prime_diff"] == 2**3072 - pyref.MUHASH_P
    # EMPTY_MUHASH (lib.rs:17-21, test_empty_hash)
    assert c_finalize(oracle, new(oracle)).hex() == golden["empty_muhash"]
    assert pyref.MuHash().finalize().hex() == golden["empty_muhash"]
    # TEST_VECTORS: test_vectors_hash, test_vectors_add_remove
    cum_c, cum_p = new(oracle), pyref.MuHash()
    for v in golden["test_vectors"]:
        d = bytes.fromhex(v["data"])
        m = new(oracle)
        oracle.ok_muhash_add_element(ctypes.byref(m), d, len(d))
        assert c_finalize(oracle, m).hex() == v["multiset_hash"]
        p = pyref.MuHash(); p.add_element(d)
        assert p.finalize().hex() == v["multiset_hash"]
        oracle.ok_muhash_add_element(ctypes.byref(cum_c), d, len(d))
        cum_p.add_element(d)
        assert c_finalize(oracle, cum_c).hex() == v["cumulative_hash"]
        assert cum_p.finalize().hex() == v["cumulative_hash"]
    for i in reversed(range(3)):
        d = bytes.fromhex(golden["test_vectors"][i]["data"])
        oracle.ok_muhash_remove_element(ctypes.byref(cum_c), d, len(d))
        cum_p.remove_element(d)
        want = golden["test_vectors"][i - 1]["cumulative_hash"] if i else golden["empty_muhash"]
        assert c_finalize(oracle, cum_c).hex() == want and cum_p.finalize().hex() == want
    # test_new_pre_computed
    pc = golden["pre_computed"]
    m, p = new(oracle), pyref.MuHash()
    for h in pc["add"]:
        oracle.ok_muhash_add_element(ctypes.byref(m), bytes.fromhex(h), 32); p.add_element(bytes.fromhex(h))
    for h in pc["remove"]:
        oracle.ok_muhash_remove_element(ctypes.byref(m), bytes.fromhex(h), 32); p.remove_element(bytes.fromhex(h))
    assert c_finalize(oracle, m).hex() == pc["finalized"] == p.finalize().hex()
    # test_serialize
    se = golden["serialize"]
    m, p = new(oracle), pyref.MuHash()
    for h in se["add"]:
        oracle.ok_muhash_add_element(ctypes.byref(m), bytes.fromhex(h), 32); p.add_element(bytes.fromhex(h))
    out = ctypes.create_string_buffer(384)
    oracle.ok_muhash_serialize(ctypes.byref(m), out)
    assert out.raw.hex() == se["serialized"] == p.serialize().hex()
    m2 = OkMuHash()
    assert oracle.ok_muhash_deserialize(ctypes.byref(m2), out.raw) == 0
    assert c_finalize(oracle, m2) == c_finalize(oracle, m)
    # test_parse_muhash_fail
    pf = golden["parse_fail"]
    assert oracle.ok_muhash_deserialize(ctypes.byref(m2), bytes.fromhex(pf["overflow"])) == -1
    assert oracle.ok_muhash_deserialize(ctypes.byref(m2), bytes.fromhex(pf["ok"])) == 0
    assert oracle.ok_muhash_deserialize(ctypes.byref(m2), b"\xff" * 384) == -1


def test_commutativity_combine_and_random_cross_check(oracle):
    rnd = random.Random(5)
    items = [bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 150))) for _ in range(24)]
    a, b, p = new(oracle), new(oracle), pyref.MuHash()
    for i, d in enumerate(items):
        tgt = a if i % 2 else b
        if i % 3 == 0:
            oracle.ok_muhash_remove_element(ctypes.byref(tgt), d, len(d)); p.remove_element(d)
        else:
            oracle.ok_muhash_add_element(ctypes.byref(tgt), d, len(d)); p.add_element(d)
    oracle.ok_muhash_combine(ctypes.byref(a), ctypes.byref(b))
    num, den = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
    oracle.ok_muhash_raw(ctypes.byref(a), num, den)
    assert int.from_bytes(num.raw, "little") == p.num and int.from_bytes(den.raw, "little") == p.den
    assert c_finalize(oracle, a) == p.finalize()
    # removing everything that was added and vice versa returns to the empty hash (test_muhash_add_remove)
    for i, d in enumerate(items):
        if i % 3 == 0:
            oracle.ok_muhash_add_element(ctypes.byref(a), d, len(d))
        else:
            oracle.ok_muhash_remove_element(ctypes.byref(a), d, len(d))
    assert c_finalize(oracle, a) == pyref.MuHash().finalize()


def test_u3072_edge_values(oracle):
    """elements at and around the modulus: p == 0, p+1 == 1, 2^3072-1 == PRIME_DIFF-1 (u3072.rs:456-492 overflow handling)"""
    P = pyref.MUHASH_P
    m2 = OkMuHash()
    for v in [P - 1, 1, 2, 2**3071, P - 2, 2**1536 + 12345]:
        assert oracle.ok_muhash_deserialize(ctypes.byref(m2), v.to_bytes(384, "little")) == 0
        other = OkMuHash()
        oracle.ok_muhash_deserialize(ctypes.byref(other), ((v * 7 + 3) % P).to_bytes(384, "little"))
        oracle.ok_muhash_combine(ctypes.byref(m2), ctypes.byref(other))
        num, den = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
        oracle.ok_muhash_raw(ctypes.byref(m2), num, den)
        assert int.from_bytes(num.raw, "little") == v * ((v * 7 + 3) % P) % P


def test_transaction_level(oracle):
    """ok_muhash_add_transaction / ok_muhash_accepted against the Python twin driven by write_utxo bytes"""
    from rusty_kaspa_b200 import simgen
    from rusty_kaspa_b200.txbatch import build_batch
    import oracle_tx
    fk, fe, txs = simgen.funded_window(12, n_keys=16, n_nonces=16, mix=(0.5, 0.2, 0.15, 0.15))
    ents, k = [], 0
    for t in txs:
        ents.append(fe[k:k + len(t["inputs"])]); k += len(t["inputs"])
    b = build_batch(txs, ents)
    ob = oracle_tx.ok_batch(b)
    accept = np.ones(len(txs), dtype=np.uint8); accept[3] = 0
    m = OkMuHash()
    oracle.ok_muhash_accepted(ctypes.byref(m), ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), accept.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(77))
    p = pyref.MuHash()
    for i, (t, es) in enumerate(zip(txs, ents)):
        if not accept[i]:
            continue
        tid = simgen.tx_id(t)
        for inp, e in zip(t["inputs"], es):
            p.remove_element(pyref.utxo_element_bytes(inp["txid"], inp["index"], e["block_daa_score"], e["amount"], e["is_coinbase"], e["spk_version"], e["script"]))
        for j, o in enumerate(t["outputs"]):
            p.add_element(pyref.utxo_element_bytes(tid, j, 77, o["value"], False, o["spk_version"], o["script"]))
    assert c_finalize(oracle, m) == p.finalize()


def test_real_utxo_commitments_and_accepted_id_roots_of_the_simpa_dag(oracle):
    """The header fields utxoCommitment and acceptedIdMerkleRoot of the reference's simpa DAG fixture, for the 210 blocks whose past
    holds only coinbase transactions (no GHOSTDAG needed, see golden_util): written by the reference itself, reproduced by the
    oracle (coinbase tx id, write_utxo with is_coinbase / block_daa_score, element expansion, product, finalize; merkle_hash)."""
    import functools
    from golden_util import simpa_dag_coinbase_only_chain_info
    by, eligible, sp = simpa_dag_coinbase_only_chain_info()
    assert len(eligible) >= 200

    @functools.lru_cache(None)
    def muhash_of(h):
        """MuHash state (numerator bytes) of block h's past UTXO set, built incrementally along the selected chain"""
        s = sp(h)
        m = OkMuHash()
        if s is None:
            oracle.ok_muhash_init(ctypes.byref(m))
            return bytes(m.num)
        assert oracle.ok_muhash_deserialize(ctypes.byref(m), muhash_of(s)) == 0
        cb = by[s]["txs"][0]
        tid = pyref.tx_id(cb)
        for i, o in enumerate(cb["outputs"]):
            d = pyref.utxo_element_bytes(tid, i, by[h]["daa_score"], o["value"], True, o["spk_version"], o["script"])
            oracle.ok_muhash_add_element(ctypes.byref(m), d, len(d))
        return bytes(m.num)

    n_nonempty = 0
    for h in eligible:
        m = OkMuHash()
        oracle.ok_muhash_deserialize(ctypes.byref(m), muhash_of(h))
        assert c_finalize(oracle, m).hex() == by[h]["utxo_commitment"], h
        n_nonempty += by[h]["utxo_commitment"] != "544eb3142c000f0ad2c76ac41f4222abbababed830eeafee4b6dc56b52d5cac0"
        s = sp(h)
        if s is not None:
            root = ctypes.create_string_buffer(32)
            oracle.ok_merkle_root(pyref.tx_id(by[s]["txs"][0]), ctypes.c_size_t(1), root)
            out = ctypes.create_string_buffer(32)
            oracle.ok_blake2b_keyed(b"MerkleBranchHash", bytes.fromhex(by[s]["accepted_id_merkle_root"]) + root.raw, ctypes.c_size_t(64), out)
            assert out.raw.hex() == by[h]["accepted_id_merkle_root"], h
    assert n_nonempty >= 190
