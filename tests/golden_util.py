"""Helpers shared by the golden-vector tests: fixture loading and the tx dict <-> bytes conventions."""
import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


def tx_from_json(t):
    """hex strings -> bytes, in the dict layout oracle/pyref.py uses"""
    return {"version": t["version"],
            "inputs": [{"txid": bytes.fromhex(i["txid"]), "index": i["index"], "sigscript": bytes.fromhex(i["sigscript"]),
                        "sequence": i["sequence"], "sig_op_count": i["sig_op_count"]} for i in t["inputs"]],
            "outputs": [{"value": o["value"], "spk_version": o["spk_version"], "script": bytes.fromhex(o["script"])} for o in t["outputs"]],
            "lock_time": t["lock_time"], "subnetwork_id": bytes.fromhex(t["subnetwork_id"]), "gas": t["gas"],
            "payload": bytes.fromhex(t["payload"]), "mass": t.get("mass", 0)}


def entry_from_json(e):
    return {"amount": e["amount"], "spk_version": e["spk_version"], "script": bytes.fromhex(e["script"]),
            "block_daa_score": e.get("block_daa_score", 0), "is_coinbase": e.get("is_coinbase", False)}


def apply_sighash_action(tx, entries, action, arg):
    """consensus/core/src/hashing/sighash.rs:655-679"""
    if action == "Output":
        tx["outputs"][arg]["value"] = 100
    elif action == "Input":
        tx["inputs"][arg]["index"] = 2
    elif action == "AmountSpent":
        entries[arg]["amount"] = 666
    elif action == "PrevScriptPublicKey":
        entries[arg]["script"] = entries[arg]["script"] + bytes([1, 2, 3])
    elif action == "Sequence":
        tx["inputs"][arg]["sequence"] = 12345
    elif action == "Payload":
        tx["payload"] = bytes([6, 6, 6, 4, 2, 0, 1, 3, 3, 7])
    elif action == "Gas":
        tx["gas"] = 1234
    elif action == "SubnetworkId":
        tx["subnetwork_id"] = bytes([6, 6, 6, 4, 2, 0, 1, 3, 3, 7]) + bytes(10)
    else:
        assert action == "NoAction"


def simpa_dag_coinbase_only_chain_info():
    """From the simpa DAG fixture: for every block whose PAST contains only coinbase transactions, the data that determines its
    header commitments without running GHOSTDAG (selected parent = max (blue_work, hash) among the level-0 parents,
    consensus/src/processes/ghostdag/ordering.rs:38-42):
      utxo_commitment(B)          = MuHash of { outputs of coinbase(C) with block_daa_score = daa_score(child of C on B's selected chain),
                                    is_coinbase = true : C on the selected chain of B, C != B }   (utxo_validation.rs:117-121,189)
      accepted_id_merkle_root(B)  = merkle_hash(accepted_id_merkle_root(SP(B)), calc_merkle_root([id(coinbase(SP(B)))]))  (KIP-15, :401-410)
    Returns (blocks_by_hash, eligible hashes in file order, selected_parent function)."""
    import functools
    import sys
    fx = load("simpa_goref_1060.json.gz")
    by = {}
    for b in fx["blocks"]:
        by[b["hash"]] = dict(b, blue_work_int=int(b["blue_work"], 16), txs=[tx_from_json(t) for t in b["transactions"]])

    def sp(h):
        ps = by[h]["parents"]
        return max(ps, key=lambda p: (by[p]["blue_work_int"], bytes.fromhex(p))) if ps else None

    sys.setrecursionlimit(10000)

    @functools.lru_cache(None)
    def past_has_spends(h):
        return any(len(by[p]["txs"]) > 1 or past_has_spends(p) for p in by[h]["parents"])

    eligible = [b["hash"] for b in fx["blocks"] if not past_has_spends(b["hash"])]
    return by, eligible, sp


def simpa_dag_replay_plan(fixture="simpa_goref_1060.json.gz"):
    """The reference's own acceptance order for a simpa DAG fixture, derived from header data only (no GHOSTDAG run):
    selected parent = max (blue_work, hash) among the level-0 parents; mergeset(B) = past(B) - past(SP) - {SP}; consensus order =
    SP first, then the rest ascending by (blue_work, hash) (processes/ghostdag/ordering.rs:38-42, utxo_validation.rs:110-160).
    Returns (fixture, by_hash, file_order, sp(h), ordered_mergeset(h), virtual_chain)."""
    fx = load(fixture)
    by, order, idx = {}, [], {}
    for n, b in enumerate(fx["blocks"]):
        by[b["hash"]] = dict(b, bw=int(b["blue_work"], 16), txs=[tx_from_json(t) for t in b["transactions"]])
        order.append(b["hash"])
        idx[b["hash"]] = n
    key = lambda h: (by[h]["bw"], bytes.fromhex(h))
    past = {}  # ancestor sets as integer bitsets over the file (topological) order
    for h in order:
        s = 0
        for p in by[h]["parents"]:
            s |= past[p] | (1 << idx[p])
        past[h] = s

    def sp(h):
        return max(by[h]["parents"], key=key) if by[h]["parents"] else None

    def ordered_mergeset(h):
        s = sp(h)
        if s is None:
            return []
        bits = past[h] & ~past[s] & ~(1 << idx[s])
        rest = []
        while bits:
            low = bits & -bits
            rest.append(order[low.bit_length() - 1])
            bits ^= low
        return [s] + sorted(rest, key=key)

    tip = max(order, key=key)
    chain = [tip]
    while sp(chain[-1]) is not None:
        chain.append(sp(chain[-1]))
    chain.reverse()
    return fx, by, order, sp, ordered_mergeset, chain
