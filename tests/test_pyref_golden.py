"""The Python twin (oracle/pyref.py) against the reference's own vectors (tests/golden, see make_golden.py)."""
import copy
import hashlib

import pyref
from golden_util import apply_sighash_action, entry_from_json, load, tx_from_json


def test_hashers_incremental():
    g = load("hashers.json")
    inputs = [bytes.fromhex(h) for h in g["inputs_hex"]]
    blake = {"TransactionHash": b"TransactionHash", "TransactionID": b"TransactionID", "TransactionSigningHash": b"TransactionSigningHash",
             "BlockHash": b"BlockHash", "MerkleBranchHash": b"MerkleBranchHash"}
    seen = 0
    for v in g["vectors"]:
        acc = b""
        for data, exp in zip(inputs, v["expected"]):
            acc += data
            if v["hasher"] in blake:
                assert pyref.blake2b_keyed(blake[v["hasher"]], acc).hex() == exp
                seen += 1
            elif v["hasher"] == "TransactionSigningHashECDSA":
                assert pyref.sha256_domain(b"TransactionSigningHashECDSA", acc).hex() == exp
                seen += 1
    assert seen >= 30


def test_tx_id_and_hash():
    for v in load("tx_hashing.json")["vectors"]:
        tx = tx_from_json(v["tx"])
        assert pyref.tx_id(tx).hex() == v["expected_id"]
        assert pyref.tx_hash(tx).hex() == v["expected_hash"]


def test_sighash_vectors():
    g = load("sighash.json")
    for v in g["vectors"]:
        tx = tx_from_json(g[v["tx"]])
        entries = [entry_from_json(e) for e in g["entries"]]
        apply_sighash_action(tx, entries, v["action"], v["action_arg"])
        assert pyref.sighash_schnorr(tx, entries, v["input_index"], v["hash_type"]).hex() == v["expected"], v["name"]


def test_mainnet_p2pk_schnorr_kat():
    c = load("check_scripts_kat.json")["cases"][0]
    tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
    ss = tx["inputs"][0]["sigscript"]
    assert ss[0] == 0x41 and len(ss) == 66 and entries[0]["script"][0] == 0x20 and entries[0]["script"][-1] == 0xAC
    msg = pyref.sighash_schnorr(tx, entries, 0, ss[65])
    assert pyref.schnorr_verify(entries[0]["script"][1:33], msg, ss[1:65]) == pyref.VALID
    # the incorrect-signature case: same signature against another key
    c2 = load("check_scripts_kat.json")["cases"][1]
    tx2, e2 = tx_from_json(c2["tx"]), [entry_from_json(e) for e in c2["entries"]]
    msg2 = pyref.sighash_schnorr(tx2, e2, 0, ss[65])
    assert pyref.schnorr_verify(e2[0]["script"][1:33], msg2, tx2["inputs"][0]["sigscript"][1:65]) == pyref.INVALID
