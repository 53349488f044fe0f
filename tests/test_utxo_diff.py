"""UtxoDiff algebra (host mirror rusty_kaspa_b200/utxo_diff.py) against the reference's own rule table
(consensus/core/src/utxo/utxo_diff.rs:270-568 test_utxo_diff_rules, extracted by tests/golden/make_golden.py), run exactly as that
test runs it (including the round trips), plus add_transaction on generated transactions against a plain dict model."""
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from golden_util import load  # noqa: E402
from rusty_kaspa_b200.utxo_diff import UtxoAlgebraError, UtxoDiff  # noqa: E402

OUTPOINT0 = (bytes(32), 0)


def _entry(spec):
    return {"amount": spec["amount"], "spk_version": 0, "script": b"", "block_daa_score": spec["block_daa_score"], "is_coinbase": spec["is_coinbase"]}


def _diff(d, entries):
    return UtxoDiff({OUTPOINT0: entries[str(k)] for k in d["add"]}, {OUTPOINT0: entries[str(k)] for k in d["remove"]})


def _run(fn):
    try:
        return ("ok", fn())
    except UtxoAlgebraError as e:
        return ("err", e.kind)


def test_reference_rule_table():
    g = load("utxo_diff_rules.json")
    entries = {k: _entry(v) for k, v in g["entries"].items()}
    assert len(g["tests"]) == 24
    for t in g["tests"]:
        this, other = _diff(t["this"], entries), _diff(t["other"], entries)
        kind, val = _run(lambda: this.diff_from(other))
        if "ok" in t["diff_from"]:
            assert kind == "ok" and val == _diff(t["diff_from"]["ok"], entries), t["name"]
            assert this.with_diff(val) == other, "reverse diff_from: " + t["name"]
        else:
            assert (kind, val) == ("err", t["diff_from"]["err"]), t["name"]
        kind, val = _run(lambda: this.with_diff(other))
        if "ok" in t["with_diff"]:
            assert kind == "ok" and val == _diff(t["with_diff"]["ok"], entries), t["name"]
            assert this.diff_from(val) == other, "reverse with_diff: " + t["name"]
        else:
            assert (kind, val) == ("err", t["with_diff"]["err"]), t["name"]
        assert this == _diff(t["this"], entries) and other == _diff(t["other"], entries)  # with_diff / diff_from do not mutate


def test_add_transaction_and_composition_against_a_set_model():
    """diffs built by add_transaction over consecutive blocks compose (with_diff) into the diff between the first and the last
    UTXO set, and diff_from recovers each step; double spends / double adds raise"""
    from rusty_kaspa_b200 import simgen
    dag = simgen.SimDag(seed=9, n_keys=16, n_nonces=32, coinbase_maturity=1, coinbase_outputs=4)
    base, cur = {}, {}
    total = UtxoDiff()
    prev_total = UtxoDiff()
    for step in range(12):
        txs, pov = dag.make_block(6)
        d = UtxoDiff()
        for tx in txs:
            ents = [cur.get((i["txid"], i["index"])) for i in tx["inputs"]]
            if any(e is None for e in ents):
                continue
            tid = simgen.tx_id(tx)
            d.add_transaction(tx, ents, tid, pov, is_coinbase=tx["subnetwork_id"][0] == 1)
            for i in tx["inputs"]:
                del cur[(i["txid"], i["index"])]
            for k, o in enumerate(tx["outputs"]):
                cur[(tid, k)] = {"amount": o["value"], "spk_version": o["spk_version"], "script": o["script"], "block_daa_score": pov, "is_coinbase": tx["subnetwork_id"][0] == 1}
        prev_total = total
        total = total.with_diff(d)
        assert prev_total.diff_from(total) == d
    # the composed diff is exactly (current set - base set, base set - current set)
    assert total.add == {o: e for o, e in cur.items() if o not in base} and total.remove == {}
    assert len(total.add) > 20
    # algebra errors of add_transaction
    tx = {"inputs": [{"txid": b"\x01" * 32, "index": 0}], "outputs": []}
    e = {"amount": 1, "spk_version": 0, "script": b"", "block_daa_score": 5, "is_coinbase": False}
    d = UtxoDiff()
    d.add_transaction(tx, [e], b"\x02" * 32, 9)
    with pytest.raises(UtxoAlgebraError) as ei:
        d.add_transaction(tx, [e], b"\x03" * 32, 9)
    assert ei.value.kind == "DoubleRemoveCall"


def test_cpp_mirror_runs_the_reference_rule_table(tmp_path):
    """kgv::UtxoDiff (include/kgv.hpp) through tests/cpp/utxo_diff_test.cpp on the same table (pure host code, no GPU)"""
    import subprocess
    root = os.path.dirname(HERE)
    src, exe = os.path.join(HERE, "cpp", "utxo_diff_test.cpp"), os.path.join(HERE, "cpp", "utxo_diff_test")
    deps = [src, os.path.join(root, "include", "kgv.hpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    g = load("utxo_diff_rules.json")
    f = lambda ks: "".join(str(k) for k in ks) or "-"
    inp = "\n".join("%s %s %s %s" % (f(t["this"]["add"]), f(t["this"]["remove"]), f(t["other"]["add"]), f(t["other"]["remove"])) for t in g["tests"]) + "\n"
    out = subprocess.run([exe], input=inp, capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(g["tests"])
    want = lambda r: ("ok %s %s" % (f(r["ok"]["add"]), f(r["ok"]["remove"]))) if "ok" in r else "err " + r["err"]
    for t, line in zip(g["tests"], out):
        a, b, trips = [x.strip() for x in line.split("|")]
        assert a == want(t["diff_from"]) and b == want(t["with_diff"]) and trips == "1", (t["name"], line)
