"""GPU + host engine: non-standard scripts are declined by the fast path (KGV_TX_NEEDS_HOST_VM) and decided by the
host script engine with GPU-verified signatures (kgv_check_scripts_host).  Expected results come from running the
same engine with oracle verdicts on the CPU (itself pinned by the reference corpus in tests/test_host_vm.py)."""
import copy

import numpy as np
import pytest

from golden_util import entry_from_json, load, tx_from_json
from rusty_kaspa_b200 import Params, TransactionValidator
from rusty_kaspa_b200.simgen import SUBNET_NATIVE, SimDag, sighash_all
from rusty_kaspa_b200.txbatch import build_batch
from rusty_kaspa_b200.validator import SCRIPT_ERR_NAMES, script_execute
from test_host_vm import oracle_verdicts

pytestmark = pytest.mark.gpu


def test_reference_kats_full_path(gpu_ctx):
    tv = TransactionValidator(gpu_ctx, Params(coinbase_maturity=100, storage_mass_parameter=0))
    for c in load("check_scripts_kat.json")["cases"]:
        tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
        tx2 = copy.deepcopy(tx)
        tx2["inputs"].append(copy.deepcopy(tx2["inputs"][-1]))
        b = build_batch([tx, tx2], [entries, entries + [copy.deepcopy(entries[-1])]])
        res = tv.validate_populated_transactions(b, entries[0]["block_daa_score"] + 1000, flags=2, host_vm=True)
        for r, exp in zip(res, (c["expected"], c["expected_duplicated_input"])):
            assert r["status"] != 11
            name = "Ok" if r["status"] == 0 else SCRIPT_ERR_NAMES[int(r["script_err"])]
            if exp == "AnyError":
                assert r["status"] == 9
            else:
                assert name == exp, (c["name"], name, exp)


def _custom_spends(n, seed):
    """non-standard but meaningful scripts: CHECKSIGVERIFY+OP_1, IF/ELSE two-key, bare 2-of-3 multisig, CLTV-guarded key"""
    dag = SimDag(seed=seed, n_keys=32, n_nonces=64)
    rng = np.random.default_rng(seed)
    txs, ents = [], []
    for t in range(n):
        kind = int(rng.integers(0, 4))
        ks = [int(x) for x in rng.choice(dag.keys.count, size=3, replace=False)]
        pk = [dag.keys.xs[k] for k in ks]
        lock_time = 0
        if kind == 0:
            spk = bytes([0x20]) + pk[0] + bytes([0xAD, 0x51])
        elif kind == 1:
            spk = bytes([0x63, 0x20]) + pk[0] + bytes([0xAC, 0x67, 0x20]) + pk[1] + bytes([0xAC, 0x68])
        elif kind == 2:
            spk = bytes([0x52]) + b"".join(bytes([0x20]) + p for p in pk) + bytes([0x53, 0xAE])
        else:
            spk = bytes([0x01, 0x64, 0xB0, 0x75, 0x20]) + pk[0] + bytes([0xAC])  # <100> CLTV DROP <pk> CHECKSIG
            lock_time = int(rng.choice([50, 100, 200]))
        entry = {"amount": 10**9, "spk_version": 0, "script": spk, "block_daa_score": 5, "is_coinbase": False}
        tx = {"version": 0, "inputs": [{"txid": bytes(rng.integers(0, 256, 32, dtype=np.uint8)), "index": 0, "sigscript": b"", "sequence": 0, "sig_op_count": 3}],
              "outputs": [{"value": 10**9 - 1, "spk_version": 0, "script": bytes([0x20]) + pk[2] + bytes([0xAC])}], "lock_time": lock_time,
              "subnetwork_id": SUBNET_NATIVE, "gas": 0, "payload": b"", "mass": 0}
        msg = sighash_all(tx, [entry], 0, False)
        bad = rng.random() < 0.25
        push = lambda k: bytes([0x41]) + (dag._sign(k, msg if not bad else bytes(32), False)) + bytes([0x01])
        if kind == 0 or kind == 3:
            ss = push(ks[0])
        elif kind == 1:
            branch = int(rng.integers(0, 2))
            ss = push(ks[0] if branch else ks[1]) + (bytes([0x51]) if branch else bytes([0x00]))
        else:
            pair = sorted(int(x) for x in rng.choice(3, size=2, replace=False))
            ss = b"".join(push(ks[p]) for p in pair)
        tx["inputs"][0]["sigscript"] = ss
        txs.append(tx)
        ents.append([entry])
    return txs, ents


def test_nonstandard_scripts_via_host_engine(gpu_ctx, oracle):
    txs, ents = _custom_spends(160, seed=3)
    b = build_batch(txs, ents)
    tv = TransactionValidator(gpu_ctx, Params(coinbase_maturity=0, storage_mass_parameter=0))
    fast = tv.validate_populated_transactions(b, 1000, flags=2)
    assert (fast["status"] == 11).all()  # every one of these shapes is declined by the fast path
    res = tv.validate_populated_transactions(b, 1000, flags=2, host_vm=True)
    names = set()
    for i in range(len(txs)):
        exp = script_execute(b, i, 0, oracle_verdicts(oracle, b))
        got = 0 if res[i]["status"] == 0 else int(res[i]["script_err"])
        assert got == exp, (i, SCRIPT_ERR_NAMES[got], SCRIPT_ERR_NAMES[exp])
        assert res[i]["status"] in (0, 9)
        names.add(SCRIPT_ERR_NAMES[got])
    assert {"Ok", "EvalFalse", "VerifyError", "NullFail", "UnsatisfiedLockTime"} <= names, names


def _load_entries(us, txs, ents):
    """puts the entries spent by `txs` into the GPU UTXO set"""
    from rusty_kaspa_b200.simgen import entries_to_arrays
    keys = np.frombuffer(b"".join(i["txid"] + int(i["index"]).to_bytes(4, "little") for t in txs for i in t["inputs"]), dtype=np.uint8).reshape(-1, 36)
    arr, arena = entries_to_arrays([e for es in ents for e in es])
    us.apply_diff(add_keys36=keys, add_entries=arr, add_bytes=arena)


def test_nonstandard_spends_through_the_table_path(gpu_ctx, oracle):
    """utxo_validation.rs:282-309 accepts any transaction whose scripts execute successfully: kgv_validate_txs (table-backed) must
    decide non-standard spends itself (never KGV_TX_NEEDS_HOST_VM), including script public keys far longer than a slot's inline
    68 bytes (kept in the overflow arena; the old stride-limited lookup truncated them), and
    validate_transactions_with_muhash_in_parallel must fold the valid ones into the MuHash.  Same through kgv_replay_window."""
    from rusty_kaspa_b200 import GpuUtxoSet, MuHash
    from rusty_kaspa_b200.replay import DagReplayer, REPLAY_ACCEPT_COINBASE
    from rusty_kaspa_b200.simgen import SUBNET_COINBASE
    txs, ents = _custom_spends(96, seed=9)
    # a 400-byte script public key: <pk> CHECKSIGVERIFY, 30 x (push 10 bytes, DROP), OP_1
    dag = SimDag(seed=77, n_keys=8, n_nonces=16)
    for j in range(6):
        pk = dag.keys.xs[j]
        spk = bytes([0x20]) + pk + bytes([0xAD]) + b"".join(bytes([0x0A]) + bytes([j + 1] * 10) + bytes([0x75]) for _ in range(30)) + bytes([0x51])
        assert len(spk) > 300
        entry = {"amount": 10**9, "spk_version": 0, "script": spk, "block_daa_score": 5, "is_coinbase": False}
        tx = {"version": 0, "inputs": [{"txid": bytes([0xC0 + j]) * 32, "index": j, "sigscript": b"", "sequence": 0, "sig_op_count": 1}],
              "outputs": [{"value": 10**9 - 1, "spk_version": 0, "script": bytes([0x20]) + pk + bytes([0xAC])}], "lock_time": 0,
              "subnetwork_id": SUBNET_NATIVE, "gas": 0, "payload": b"", "mass": 0}
        msg = sighash_all(tx, [entry], 0, False)
        sig = dag._sign(j, msg if j != 3 else bytes(32), False)  # one of them carries a wrong signature
        tx["inputs"][0]["sigscript"] = bytes([0x41]) + sig + bytes([0x01])
        txs.append(tx); ents.append([entry])
    pb = build_batch(txs, ents)
    exp = [script_execute(pb, i, 0, oracle_verdicts(oracle, pb)) for i in range(len(txs))]
    assert exp[-6:] == [0, 0, 0, 25, 0, 0] and 0 in exp[:96] and any(e != 0 for e in exp[:96])
    tv = TransactionValidator(gpu_ctx, Params(coinbase_maturity=0, storage_mass_parameter=0))
    us = GpuUtxoSet(gpu_ctx, 1 << 12)
    _load_entries(us, txs, ents)
    b = build_batch(txs)
    res, mu = tv.validate_transactions_with_muhash_in_parallel(us, b, 1000, flags=2)
    assert not (res["status"] == 11).any()
    got = [0 if r["status"] == 0 else int(r["script_err"]) for r in res]
    assert got == exp
    acc = (res["status"] == 0).astype(np.uint8)
    assert mu.numerator == MuHash.from_transactions(gpu_ctx, pb, acc, 1000).numerator and acc[-1] == 1 and acc[-3] == 0
    us.close()
    # the same spends as one block of a replay window (tx 0 = coinbase)
    cb = {"version": 0, "inputs": [], "outputs": [{"value": 5, "spk_version": 0, "script": bytes([0x51])}], "lock_time": 0, "subnetwork_id": SUBNET_COINBASE,
          "gas": 0, "payload": b"x", "mass": 0}
    r = DagReplayer(gpu_ctx, Params(coinbase_maturity=0, storage_mass_parameter=0), 1 << 12)
    _load_entries(r.us, txs, ents)
    for t in txs:  # mass check is on in the replay: commit the storage mass the context rules expect (C = 0 -> 0)
        t["mass"] = 0
    out = r.replay_windowed([([cb] + txs, 1000, REPLAY_ACCEPT_COINBASE)])[0]
    assert r.last_stats["n_host_vm"] == len(txs)
    assert [0 if x["status"] == 0 else int(x["script_err"]) for x in out[1:]] == exp and out[0]["status"] == 12
    assert r.us.count() == 1 + int(acc.sum()) + (len(txs) - int(acc.sum()))  # coinbase output + one output per accepted tx + unspent entries of rejected txs
    r.close()
