"""Exploration: replay the whole simpa DAG fixture with the CPU oracle and compare every header's utxoCommitment."""
import sys, os, ctypes, functools, copy, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pyref, oracle_tx
from golden_util import load, tx_from_json
from rusty_kaspa_b200.txbatch import build_batch

class OkMuHash(ctypes.Structure):
    _fields_ = [("num", ctypes.c_uint64 * 48), ("den", ctypes.c_uint64 * 48)]

O = ctypes.CDLL(os.path.join(ROOT, "oracle", "libkaspa_oracle.so")); O.ok_secp_init()
fx = load("simpa_goref_1060.json.gz")
by = {}
order = []
for b in fx["blocks"]:
    by[b["hash"]] = dict(b, bw=int(b["blue_work"], 16), txs=[tx_from_json(t) for t in b["transactions"]])
    order.append(b["hash"])
key = lambda h: (by[h]["bw"], bytes.fromhex(h))
sp = lambda h: max(by[h]["parents"], key=key) if by[h]["parents"] else None
past = {}
for h in order:  # file order is topological
    s = set()
    for p in by[h]["parents"]:
        s.add(p); s |= past[p]
    past[h] = s
prm = oracle_tx.params(coinbase_maturity=fx["coinbase_maturity"] or 200, storage_mass_parameter=fx["storage_mass_parameter"])
print("maturity", fx["coinbase_maturity"], "C", fx["storage_mass_parameter"])
state, mh = {}, {}
ok = bad = 0
n_acc = n_rej = 0
t0 = time.time()
for h in order:
    b = by[h]; s = sp(h)
    m = OkMuHash()
    if s is None:
        O.ok_muhash_init(ctypes.byref(m)); st = {}
    else:
        ctypes.memmove(ctypes.byref(m), ctypes.byref(mh[s]), ctypes.sizeof(m)); st = dict(state[s])
        pov = b["daa_score"]
        def add(txid, i, o, coinbase):
            e = {"amount": o["value"], "spk_version": o["spk_version"], "script": o["script"], "block_daa_score": pov, "is_coinbase": coinbase}
            st[(txid, i)] = e
            d = pyref.utxo_element_bytes(txid, i, pov, o["value"], coinbase, o["spk_version"], o["script"])
            O.ok_muhash_add_element(ctypes.byref(m), d, len(d))
        cb = by[s]["txs"][0]; cid = pyref.tx_id(cb)
        for i, o in enumerate(cb["outputs"]): add(cid, i, o, True)
        ms = [s] + sorted(past[h] - past[s] - {s}, key=key)
        for mb in ms:
            for tx in by[mb]["txs"][1:]:
                ents = [st.get((i["txid"], i["index"])) for i in tx["inputs"]]
                if any(e is None for e in ents): n_rej += 1; continue
                bt = build_batch([tx], [ents])
                r = oracle_tx.validate_populated(O, bt, 0, pov, 0, prm)
                if int(r["status"]) != 0: n_rej += 1; continue
                n_acc += 1
                for i, e in zip(tx["inputs"], ents):
                    del st[(i["txid"], i["index"])]
                    d = pyref.utxo_element_bytes(i["txid"], i["index"], e["block_daa_score"], e["amount"], e["is_coinbase"], e["spk_version"], e["script"])
                    O.ok_muhash_remove_element(ctypes.byref(m), d, len(d))
                tid = pyref.tx_id(tx)
                for i, o in enumerate(tx["outputs"]): add(tid, i, o, False)
    state[h], mh[h] = st, m
    mm = OkMuHash(); ctypes.memmove(ctypes.byref(mm), ctypes.byref(m), ctypes.sizeof(m))
    out = ctypes.create_string_buffer(32); O.ok_muhash_finalize(ctypes.byref(mm), out)
    if out.raw.hex() == b["utxo_commitment"]: ok += 1
    else:
        bad += 1
        if bad <= 3: print("mismatch at", h[:8], "daa", b["daa_score"], "mergeset", len(past[h] - past[s] - {s}) + 1 if s else 0)
print("commitments ok", ok, "bad", bad, "accepted txs (sum over povs)", n_acc, "rejected", n_rej, "%.1fs" % (time.time() - t0))
