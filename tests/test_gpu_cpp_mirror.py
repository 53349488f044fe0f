"""The C++ host mirror (include/kgv.hpp) driven by tests/cpp/host_mirror_test.cpp on dumped data: every printed result is
compared with the oracle or with the reference's known answers (MuHash)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_tx  # noqa: E402

pytestmark = pytest.mark.gpu
BIN = os.path.join(HERE, "cpp", "host_mirror_test")


def _build():
    src = BIN + ".cpp"
    deps = [src, os.path.join(ROOT, "include", "kgv.hpp"), os.path.join(ROOT, "include", "kgv.h")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", BIN, src, "-L" + os.path.join(ROOT, "rusty_kaspa_b200"), "-l:libkgv.so",
                        "-Wl,-rpath,$ORIGIN/../../rusty_kaspa_b200"], check=True)


class OkMuHash(ctypes.Structure):
    _fields_ = [("num", ctypes.c_uint64 * 48), ("den", ctypes.c_uint64 * 48)]


def test_cpp_mirror_end_to_end(tmp_path, oracle):
    from rusty_kaspa_b200 import simgen, workload as W
    from rusty_kaspa_b200.txbatch import build_batch
    _build()
    golden = json.load(open(os.path.join(HERE, "golden", "muhash.json")))
    d = str(tmp_path)
    # transactions: a mixed window, one corrupted signature, one missing entry
    fk, fe, txs = simgen.funded_window(60, n_keys=16, n_nonces=16, mix=(0.5, 0.2, 0.15, 0.15))
    ents, k = [], 0
    for t in txs:
        ents.append(list(fe[k:k + len(t["inputs"])])); k += len(t["inputs"])
    ss = txs[7]["inputs"][0]["sigscript"]
    txs[7]["inputs"][0]["sigscript"] = ss[:20] + bytes([ss[20] ^ 4]) + ss[21:]
    b = build_batch(txs, ents)
    for name, arr in (("txs", b.txs), ("inputs", b.inputs), ("outputs", b.outputs), ("entries", b.entries), ("arena", b.arena)):
        arr.tofile(os.path.join(d, name + ".bin"))
    ae, ab = simgen.entries_to_arrays(fe)
    fk.tofile(os.path.join(d, "fund_keys.bin")); ae.tofile(os.path.join(d, "fund_entries.bin")); ab.tofile(os.path.join(d, "fund_arena.bin"))
    first = np.array([0, 1, 1, 20, 45, 60], dtype=np.uint32)
    first.tofile(os.path.join(d, "blocks.bin"))
    pc = golden["pre_computed"]
    with open(os.path.join(d, "elements.txt"), "w") as f:
        f.write("\n".join(pc["add"] + ["-" + h for h in pc["remove"]]) + "\n")
    pk, msg, sig, kind = W.schnorr_triples(64, seed=3, n_keys=16, n_nonces=16, frac_bitflip=0.2, frac_adversarial=0.2)
    np.concatenate([pk.reshape(-1, 32), msg.reshape(-1, 32), sig.reshape(-1, 64)], axis=1).tofile(os.path.join(d, "triples.bin"))
    out = subprocess.run([BIN, d, str(simgen.DEFAULT_STORAGE_MASS_PARAMETER)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = {l.split(" ", 1)[0]: l.split(" ", 1)[1] if " " in l else "" for l in out.stdout.strip().splitlines()}
    # MuHash known answers of the reference
    assert lines["muhash_empty"] == golden["empty_muhash"] and lines["muhash_elements"] == pc["finalized"]
    # signatures vs oracle
    exp = np.zeros(64, dtype=np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    oracle.ok_schnorr_verify_batch(vp(pk), vp(msg), vp(sig), ctypes.c_size_t(64), vp(exp), 2)
    assert [int(x) for x in lines["schnorr"].split()] == exp.tolist()
    # populated validation vs oracle
    op = oracle_tx.params(coinbase_maturity=100, storage_mass_parameter=simgen.DEFAULT_STORAGE_MASS_PARAMETER)
    want = []
    for i in range(len(txs)):
        e = oracle_tx.validate_populated(oracle, b, i, 10, 0, op)
        want.append("%d:%d:%d" % (int(e["status"]), int(e["script_err"]), int(e["fee"]) if int(e["status"]) == 0 else 0))
    got = [g if g.split(":")[0] == "0" else ":".join(g.split(":")[:2] + ["0"]) for g in lines["populated"].split()]
    assert got == want and want[7].split(":")[0] != "0"
    assert int(lines["utxo_count"]) == len(fe)
    st = [int(x) for x in lines["in_parallel"].split()]
    assert st == [int(w.split(":")[0]) for w in want]
    # MuHash of the accepted txs vs oracle
    accept = (np.array(st) == 0).astype(np.uint8)
    m = OkMuHash()
    ob = oracle_tx.ok_batch(b)
    oracle.ok_muhash_accepted(ctypes.byref(m), ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), vp(accept), ctypes.c_uint64(10))
    num, den = ctypes.create_string_buffer(384), ctypes.create_string_buffer(384)
    oracle.ok_muhash_raw(ctypes.byref(m), num, den)
    assert lines["tx_muhash_num"] == num.raw.hex() and lines["tx_muhash_den"] == den.raw.hex()
    assert lines["commitment_matches"] == "1"
    # pruning-point import through the mirror (UtxoSet::iterator -> append_imported_pruning_point_utxos in chunks of 7): same size, same commitment
    n_src, n_dst, ms_ok, set_ok = lines["pruning_import"].split()
    assert n_src == n_dst and int(n_src) > 40 and ms_ok == "1" and set_ok == "1"
    # composed view + SigCache through the C++ mirror: same verdicts, base untouched, second pass served by the cache, spent-in-view outpoints missing
    same, base_count, second_hits, first_inserts, second_inserts, missing, n_acc = (int(x) for x in lines["view"].split())
    assert same == 1 and base_count == len(fe) and first_inserts > 40 and second_hits >= first_inserts and second_inserts <= 2 and missing == n_acc > 40
    # block bodies vs oracle
    roots, bodies = lines["merkle"].split(), lines["bodies"].split()
    hashes = ctypes.create_string_buffer(32 * len(txs))
    oracle.ok_tx_hashes(ctypes.byref(ob), hashes, 1)
    for g in range(len(first) - 1):
        o32 = ctypes.create_string_buffer(32)
        oracle.ok_merkle_root(hashes.raw[32 * int(first[g]):32 * int(first[g + 1])], ctypes.c_size_t(int(first[g + 1] - first[g])), o32)
        assert roots[g] == o32.raw.hex()
        idx = ctypes.c_uint32()
        stt = oracle.ok_block_set_checks(ctypes.byref(ob), ctypes.c_uint32(int(first[g])), ctypes.c_uint32(int(first[g + 1])), ctypes.byref(idx))
        assert bodies[g] == "%d:%d" % (stt, idx.value if stt else 0)
    assert lines["size_mismatch"].startswith("throws")
