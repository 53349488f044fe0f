"""Flat SoA transaction batches — the memory layout that crosses the C ABI (include/kgv.h:
kgv_tx / kgv_input / kgv_output / kgv_utxo_entry + one byte arena).

Host-side mirror of the reference data model (consensus/core/src/tx.rs: Transaction :165-185,
TransactionInput :93-101, TransactionOutpoint :72-77, TransactionOutput :121-125, UtxoEntry :49-57,
ScriptPublicKey tx/script_public_key.rs:22-25).  A Rust shim fills the same arrays from `&Transaction`
(INTEGRATION.md); here they are built from plain dicts:
    tx    = {version, inputs:[{txid(32B), index, sigscript, sequence, sig_op_count}],
             outputs:[{value, spk_version, script}], lock_time, subnetwork_id(20B), gas, payload, mass}
    entry = {amount, spk_version, script, block_daa_score, is_coinbase}
"""
import numpy as np

TX_DTYPE = np.dtype([("first_input", "<u4"), ("n_inputs", "<u4"), ("first_output", "<u4"), ("n_outputs", "<u4"),
                     ("lock_time", "<u8"), ("gas", "<u8"), ("mass", "<u8"), ("payload_off", "<u4"), ("payload_len", "<u4"),
                     ("version", "<u2"), ("subnetwork_id", "u1", (20,)), ("flags", "u1"), ("pad_", "u1")], align=False)
INPUT_DTYPE = np.dtype([("prev_txid", "u1", (32,)), ("prev_index", "<u4"), ("sigscript_off", "<u4"), ("sigscript_len", "<u4"),
                        ("sig_op_count", "u1"), ("pad_", "u1", (3,)), ("sequence", "<u8")], align=False)
OUTPUT_DTYPE = np.dtype([("value", "<u8"), ("script_off", "<u4"), ("script_len", "<u4"), ("spk_version", "<u2"), ("pad_", "u1", (6,))], align=False)
ENTRY_DTYPE = np.dtype([("amount", "<u8"), ("block_daa_score", "<u8"), ("script_off", "<u4"), ("script_len", "<u4"),
                        ("spk_version", "<u2"), ("is_coinbase", "u1"), ("pad_", "u1", (5,))], align=False)
assert TX_DTYPE.itemsize == 72 and INPUT_DTYPE.itemsize == 56 and OUTPUT_DTYPE.itemsize == 24 and ENTRY_DTYPE.itemsize == 32

SUBNETWORK_NATIVE = bytes(20)
SUBNETWORK_COINBASE = bytes([1]) + bytes(19)
TX_FLAG_COINBASE = 1


class TxBatch:
    """txs / inputs / outputs / entries structured arrays + the byte arena they point into."""

    def __init__(self, txs, inputs, outputs, entries, arena):
        self.txs, self.inputs, self.outputs, self.entries, self.arena = txs, inputs, outputs, entries, arena

    @property
    def n_txs(self):
        return len(self.txs)

    @property
    def n_inputs(self):
        return len(self.inputs)


def build_batch(txs, entries_per_tx=None):
    """txs: list of tx dicts; entries_per_tx: optional list (per tx) of lists (per input) of entry dicts or None."""
    n_in = sum(len(t["inputs"]) for t in txs)
    n_out = sum(len(t["outputs"]) for t in txs)
    T = np.zeros(len(txs), dtype=TX_DTYPE)
    I = np.zeros(n_in, dtype=INPUT_DTYPE)
    O = np.zeros(n_out, dtype=OUTPUT_DTYPE)
    E = np.zeros(n_in, dtype=ENTRY_DTYPE)
    arena = bytearray()

    def put(b):
        off = len(arena)
        arena.extend(b)
        return off, len(b)

    ii = oi = 0
    for ti, t in enumerate(txs):
        r = T[ti]
        r["first_input"], r["n_inputs"], r["first_output"], r["n_outputs"] = ii, len(t["inputs"]), oi, len(t["outputs"])
        r["lock_time"], r["gas"], r["mass"], r["version"] = t["lock_time"], t["gas"], t.get("mass", 0), t["version"]
        r["subnetwork_id"] = np.frombuffer(t["subnetwork_id"], dtype=np.uint8)
        r["flags"] = TX_FLAG_COINBASE if t["subnetwork_id"] == SUBNETWORK_COINBASE else 0
        r["payload_off"], r["payload_len"] = put(t["payload"])
        for k, i in enumerate(t["inputs"]):
            x = I[ii]
            x["prev_txid"] = np.frombuffer(i["txid"], dtype=np.uint8)
            x["prev_index"], x["sequence"], x["sig_op_count"] = i["index"], i["sequence"], i["sig_op_count"]
            x["sigscript_off"], x["sigscript_len"] = put(i["sigscript"])
            ent = entries_per_tx[ti][k] if entries_per_tx is not None and entries_per_tx[ti] is not None else None
            if ent is not None:
                e = E[ii]
                e["amount"], e["block_daa_score"], e["spk_version"] = ent["amount"], ent.get("block_daa_score", 0), ent["spk_version"]
                e["is_coinbase"] = 1 if ent.get("is_coinbase", False) else 0
                e["script_off"], e["script_len"] = put(ent["script"])
            ii += 1
        for o in t["outputs"]:
            y = O[oi]
            y["value"], y["spk_version"] = o["value"], o["spk_version"]
            y["script_off"], y["script_len"] = put(o["script"])
            oi += 1
    return TxBatch(T, I, O, E, np.frombuffer(bytes(arena) + bytes(8), dtype=np.uint8).copy())
