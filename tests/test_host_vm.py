"""The product's HOST script engine (libkgv.so: csrc/host/kgv_script_vm.cpp) against the reference's own
script corpus: the 850 rows of crypto/txscript/test-data/script_tests.json, run exactly like
test_bitcoind_tests (crypto/txscript/src/lib.rs:1366-1555), plus the mainnet KATs.  No GPU is needed:
kgv_script_execute takes a verdict provider, here backed by the CPU oracle (sighash + verify)."""
import copy

import numpy as np
import pytest

import oracle_tx
from golden_util import entry_from_json, load, tx_from_json
from rusty_kaspa_b200.txbatch import build_batch
from rusty_kaspa_b200.validator import SCRIPT_ERR_NAMES, script_execute

# result_name() of the reference harness (lib.rs:1467-1514): TxScriptError variant -> accepted expectation strings
RESULT_NAMES = {
    "Ok": ["OK"], "NumberTooBig": ["UNKNOWN_ERROR"], "Serialization": ["UNKNOWN_ERROR"], "PubKeyFormat": ["PUBKEYFORMAT"], "EvalFalse": ["EVAL_FALSE"],
    "EmptyStack": ["EMPTY_STACK", "EVAL_FALSE", "UNBALANCED_CONDITIONAL", "INVALID_ALTSTACK_OPERATION"], "NullFail": ["NULLFAIL"], "SigLength": ["NULLFAIL"],
    "InvalidSigHashType": ["SIG_HASHTYPE"], "SignatureScriptNotPushOnly": ["SIG_PUSHONLY"], "CleanStack": ["CLEANSTACK"], "OpcodeReserved": ["BAD_OPCODE"],
    "MalformedPush": ["BAD_OPCODE"], "InvalidOpcode": ["BAD_OPCODE"], "ErrUnbalancedConditional": ["UNBALANCED_CONDITIONAL"],
    "InvalidState(condition stack empty)": ["UNBALANCED_CONDITIONAL"], "EarlyReturn": ["OP_RETURN"], "VerifyError": ["VERIFY", "EQUALVERIFY"],
    "InvalidStackOperation": ["INVALID_STACK_OPERATION", "INVALID_ALTSTACK_OPERATION"], "InvalidState(pick at an invalid location)": ["INVALID_STACK_OPERATION"],
    "InvalidState(roll at an invalid location)": ["INVALID_STACK_OPERATION"], "OpcodeDisabled": ["DISABLED_OPCODE"], "ElementTooBig": ["PUSH_SIZE"],
    "TooManyOperations": ["OP_COUNT"], "StackSizeExceeded": ["STACK_SIZE"], "InvalidPubKeyCount": ["PUBKEY_COUNT"], "InvalidSignatureCount": ["SIG_COUNT"],
    "NotMinimalData": ["MINIMALDATA", "UNKNOWN_ERROR"], "UnsatisfiedLockTime": ["UNSATISFIED_LOCKTIME"], "InvalidState(expected boolean)": ["MINIMALIF"],
    "ScriptSize": ["SCRIPT_SIZE"],
}
U64_MAX = 2**64 - 1


def spending_tx(sigscript, spk):
    """create_spending_transaction (lib.rs:1366-1397)"""
    from rusty_kaspa_b200.simgen import tx_id
    coinbase = {"version": 1, "inputs": [{"txid": bytes(32), "index": 0xFFFFFFFF, "sigscript": bytes([0, 0]), "sequence": U64_MAX, "sig_op_count": 20}],
                "outputs": [{"value": 0, "spk_version": 0, "script": spk}], "lock_time": 0, "subnetwork_id": bytes(20), "gas": 0, "payload": b"", "mass": 0}
    tx = {"version": 1, "inputs": [{"txid": tx_id(coinbase), "index": 0, "sigscript": sigscript, "sequence": U64_MAX, "sig_op_count": 20}],
          "outputs": [{"value": 0, "spk_version": 0, "script": b""}], "lock_time": 0, "subnetwork_id": bytes(20), "gas": 0, "payload": b"", "mass": 0}
    entry = {"amount": 0, "spk_version": 0, "script": spk, "block_daa_score": 0, "is_coinbase": True}
    return tx, [entry]


def oracle_verdicts(oracle, batch, tx_index=0):
    def verdict(rq):
        rel = rq.input - int(batch.txs[rq.tx]["first_input"])
        msg = oracle_tx.sighash(oracle, batch, rq.tx, rel, rq.hash_type, ecdsa=bool(rq.ecdsa))
        key, sig = bytes(rq.key[:rq.key_len]), bytes(rq.sig)
        return oracle.ok_ecdsa_verify(key, msg, sig) if rq.ecdsa else oracle.ok_schnorr_verify(key, msg, sig)
    return verdict


def test_reference_script_corpus(oracle):
    rows = load("script_tests.json.gz")["rows"]
    assert len(rows) == 850
    seen, failures = set(), []
    for i, r in enumerate(rows):
        if "builder_error" in r:
            assert r["builder_error"] == "ElementExceedsMaxSize" and r["expected"] == "PUSH_SIZE"
            continue
        tx, entries = spending_tx(bytes.fromhex(r["sigscript"]), bytes.fromhex(r["spk"]))
        b = build_batch([tx], [entries])
        name = SCRIPT_ERR_NAMES[script_execute(b, 0, 0, oracle_verdicts(oracle, b))]
        seen.add(name)
        if r["expected"] not in RESULT_NAMES.get(name, []):
            failures.append((i, r["sig_text"][:60], r["spk_text"][:60], r["expected"], name))
    assert not failures, failures[:10]
    assert len(seen) >= 20  # the corpus exercises most error classes


def test_mainnet_kats_through_host_vm(oracle):
    """tx_validation_in_utxo_context.rs:228-709 — every case, including the non-standard shapes the GPU fast path declines"""
    for c in load("check_scripts_kat.json")["cases"]:
        tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
        tx2 = copy.deepcopy(tx)
        tx2["inputs"].append(copy.deepcopy(tx2["inputs"][-1]))
        for t, e, exp in ((tx, entries, c["expected"]), (tx2, entries + [copy.deepcopy(entries[-1])], c["expected_duplicated_input"])):
            b = build_batch([t], [e])
            got = "Ok"
            for k in range(len(t["inputs"])):
                err = script_execute(b, 0, k, oracle_verdicts(oracle, b))
                if err:
                    got = SCRIPT_ERR_NAMES[err]
                    break
            if exp == "AnyError":
                assert got != "Ok", c["name"]
            else:
                assert got == exp, (c["name"], got, exp)


def test_host_vm_agrees_with_oracle_on_standard_classes(oracle):
    """cross-check: the full engine and the oracle's class-based restatement must agree on generated standard spends"""
    from rusty_kaspa_b200.simgen import SimDag
    dag = SimDag(seed=21, n_keys=32, n_nonces=32, mix=(0.3, 0.2, 0.25, 0.25), frac_invalid=0.3, coinbase_maturity=0, coinbase_outputs=8)
    st = oracle_tx.State(oracle)
    p = oracle_tx.params(coinbase_maturity=0, storage_mass_parameter=dag.C)
    checked, errs = 0, set()
    for _ in range(12):
        txs, pov = dag.make_block(10)
        b = build_batch(txs)
        res = st.validate(b, pov, 0, p, threads=1)
        ents = [[st.get(i["txid"] + i["index"].to_bytes(4, "little")) for i in t["inputs"]] for t in txs]
        keep = [k for k, (t, e) in enumerate(zip(txs, ents)) if t["inputs"] and all(x is not None for x in e)]
        pb = build_batch([txs[k] for k in keep], [ents[k] for k in keep])
        for ti in range(len(keep)):
            for ii in range(len(txs[keep[ti]]["inputs"])):
                exp = oracle_tx.check_script_std(oracle, pb, ti, ii)
                got = script_execute(pb, ti, ii, oracle_verdicts(oracle, pb))
                assert got == exp, (ti, ii, SCRIPT_ERR_NAMES[got], oracle_tx.SCRIPT_ERR[exp])
                errs.add(got)
                checked += 1
        acc = np.array([1 if (i == 0 or res[i]["status"] == 0) else 0 for i in range(len(txs))], dtype=np.uint8)
        st.accept(b, acc, pov)
        st.commit()
    assert checked > 100 and len(errs) >= 4
    st.close()
