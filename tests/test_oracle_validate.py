"""Oracle validation logic against the reference's real-signature KATs (tests/golden/check_scripts_kat.json,
from tx_validation_in_utxo_context.rs:228-709) and its storage-mass cases."""
import copy

import oracle_tx
from golden_util import entry_from_json, load, tx_from_json
from rusty_kaspa_b200.txbatch import build_batch


def _run(oracle, tx, entries):
    b = build_batch([tx], [entries])
    for i in range(len(tx["inputs"])):
        err = oracle_tx.check_script_std(oracle, b, 0, i)
        if err != 0:
            return oracle_tx.SCRIPT_ERR[err]
    return "Ok"


def test_check_scripts_kats(oracle):
    seen = set()
    for c in load("check_scripts_kat.json")["cases"]:
        tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
        got = _run(oracle, tx, entries)
        tx2 = copy.deepcopy(tx)
        tx2["inputs"].append(copy.deepcopy(tx2["inputs"][-1]))
        got2 = _run(oracle, tx2, entries + [copy.deepcopy(entries[-1])])
        for g, exp in ((got, c["expected"]), (got2, c["expected_duplicated_input"])):
            if g == "NonStandard":
                # shapes outside the three standard classes are the host VM's job (tests/test_host_vm.py pins them)
                assert c["name"] in ("check_empty_incorrect_multi_signature_test", "check_non_push_only_script_sig_test"), c["name"]
            elif exp == "AnyError":
                assert g != "Ok", c["name"]
            else:
                assert g == exp, (c["name"], g, exp)
            seen.add(g)
    assert {"Ok", "EvalFalse", "NullFail"} <= seen


def test_validate_populated_fee_and_context_rules(oracle):
    c = load("check_scripts_kat.json")["cases"][0]
    tx, entries = tx_from_json(c["tx"]), [entry_from_json(e) for e in c["entries"]]
    p = oracle_tx.params(coinbase_maturity=100)
    b = build_batch([tx], [entries])
    tx["mass"] = oracle_tx.storage_mass(oracle, b, 0, p.storage_mass_parameter)
    b = build_batch([tx], [entries])
    pov = entries[0]["block_daa_score"] + 10
    r = oracle_tx.validate_populated(oracle, b, 0, pov, 0, p)
    assert r["status"] == 0 and r["fee"] == entries[0]["amount"] - sum(o["value"] for o in tx["outputs"])
    # wrong mass
    bad = dict(tx, mass=tx["mass"] + 1)
    assert oracle_tx.validate_populated(oracle, build_batch([bad], [entries]), 0, pov, 0, p)["status"] == 7
    assert oracle_tx.validate_populated(oracle, build_batch([bad], [entries]), 0, pov, 2, p)["status"] == 0  # SkipMassCheck
    # immature coinbase
    e2 = [dict(entries[0], is_coinbase=True)]
    assert oracle_tx.validate_populated(oracle, build_batch([tx], [e2]), 0, entries[0]["block_daa_score"] + 99, 0, p)["status"] == 2
    assert oracle_tx.validate_populated(oracle, build_batch([tx], [e2]), 0, entries[0]["block_daa_score"] + 100, 0, p)["status"] == 0  # matured; is_coinbase is not part of the sighash (sighash.rs:252-255)
    # spend too high
    e3 = [dict(entries[0], amount=1)]
    assert oracle_tx.validate_populated(oracle, build_batch([tx], [e3]), 0, pov, 2, p)["status"] == 5
    # sequence lock: bit 63 clear, relative lock 20 -> needs daa + 20 - 1 < pov
    t4 = copy.deepcopy(tx)
    t4["inputs"][0]["sequence"] = 20
    assert oracle_tx.validate_populated(oracle, build_batch([t4], [entries]), 0, entries[0]["block_daa_score"] + 19, 2, p)["status"] == 8
    r = oracle_tx.validate_populated(oracle, build_batch([t4], [entries]), 0, entries[0]["block_daa_score"] + 20, 2, p)
    assert r["status"] == 9 and r["script_err"] == 1  # lock satisfied; the sequence change breaks the signature (EvalFalse)
    # input amount too high
    e5 = [dict(entries[0], amount=oracle_tx.MAX_SOMPI + 1)]
    assert oracle_tx.validate_populated(oracle, build_batch([tx], [e5]), 0, pov, 2, p)["status"] == 4
    # skip script checks accepts a broken signature
    assert oracle_tx.validate_populated(oracle, build_batch([t4], [entries]), 0, entries[0]["block_daa_score"] + 20, 1, oracle_tx.params(storage_mass_parameter=0))["status"] in (0, 7)


def test_storage_mass_cases_of_the_reference(oracle):
    """consensus/core/src/mass/mod.rs:516-729 (tests/golden/storage_mass.json): the 8 explicit values of test_storage_mass and the 8 plurality pairs
    of test_storage_mass_pluralities (equal, non-zero mass) through the oracle's ok_storage_mass"""
    from golden_util import storage_mass_cases
    groups = {}
    n_exact = 0
    for name, tx, ents, C, expected, group in storage_mass_cases():
        m = oracle_tx.storage_mass(oracle, build_batch([tx], [ents]), 0, C)
        if expected is not None:
            assert m == expected, (name, m, expected)
            n_exact += 1
        else:
            groups.setdefault(group, []).append(m)
    assert n_exact == 8 and len(groups) == 8
    for g, (a, b) in groups.items():
        assert a == b and a not in (0, None), (g, a, b)
