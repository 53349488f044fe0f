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
