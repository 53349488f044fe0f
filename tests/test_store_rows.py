"""RocksDB row formats of the reference's UTXO store: kgv_utxo_rows_encode / _decode against a literal Python restatement of
bincode 1.x (fixint, little-endian, u64 lengths) over `UtxoEntry` (consensus/core/src/tx.rs:49-57) and of `UtxoKey::as_ref` (utxo_set.rs:36-44).
No reference-held serialized row exists in the tree ("parity unpinned" for the byte strings; the formats themselves are the crates' documented defaults)."""
import struct

import numpy as np
import pytest

from rusty_kaspa_b200 import KgvError
from rusty_kaspa_b200.store_rows import decode_rows, encode_rows
from rusty_kaspa_b200.txbatch import ENTRY_DTYPE


def py_key(txid, index):
    raw = txid + struct.pack("<I", index)
    idx = raw[32:]
    r = max((i for i in range(4) if idx[i] != 0), default=0)   # rposition(|v| v != 0).unwrap_or(0)
    return raw[:32 + r + 1]


def py_value(amount, version, script, daa, cb):
    return struct.pack("<QHQ", amount, version, len(script)) + script + struct.pack("<Q?", daa, cb)


def test_rows_roundtrip_and_match_the_restated_formats():
    rng = np.random.default_rng(5)
    n = 500
    idxs = [0, 1, 255, 256, 65535, 65536, 0x01000000, 0xFFFFFFFF, 0x00FF0000] + [int(x) for x in rng.integers(0, 2**32, size=n - 9)]
    scripts = [bytes(rng.integers(0, 256, size=int(l), dtype=np.uint8)) for l in rng.choice([0, 1, 34, 35, 68, 69, 300], size=n)]
    keys = np.zeros((n, 36), dtype=np.uint8)
    ent = np.zeros(n, dtype=ENTRY_DTYPE)
    arena = bytearray()
    want_k, want_v = [], []
    for i in range(n):
        txid = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
        keys[i] = np.frombuffer(txid + struct.pack("<I", idxs[i]), dtype=np.uint8)
        amount, daa, ver, cb = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**40)), int(rng.integers(0, 3)), bool(rng.integers(0, 2))
        ent[i] = (amount, daa, len(arena), len(scripts[i]), ver, 1 if cb else 0, [0] * 5)
        arena += scripts[i]
        want_k.append(py_key(txid, idxs[i])); want_v.append(py_value(amount, ver, scripts[i], daa, cb))
    kr, ko, vr, vo = encode_rows(keys, ent, np.frombuffer(bytes(arena) + bytes(8), dtype=np.uint8))
    assert [kr[int(ko[i]):int(ko[i + 1])] for i in range(n)] == want_k
    assert [vr[int(vo[i]):int(vo[i + 1])] for i in range(n)] == want_v
    assert len(want_k[0]) == 33 and len(want_k[2]) == 33 and len(want_k[3]) == 34 and len(want_k[7]) == 36 and len(want_k[8]) == 35
    assert len(want_v[[len(s) for s in scripts].index(34)]) == 61          # the survey's "~60-65 B for a standard spk"
    k2, e2, a2 = decode_rows(kr, ko, vr, vo)
    assert (k2 == keys).all()
    for f in ("amount", "block_daa_score", "script_len", "spk_version", "is_coinbase"):
        assert (e2[f] == ent[f]).all(), f
    for i in range(n):
        assert bytes(a2[int(e2[i]["script_off"]):int(e2[i]["script_off"] + e2[i]["script_len"])]) == scripts[i]


def test_malformed_rows_are_rejected():
    keys = np.zeros((1, 36), dtype=np.uint8)
    ent = np.zeros(1, dtype=ENTRY_DTYPE)
    ent[0]["script_len"] = 4
    kr, ko, vr, vo = encode_rows(keys, ent, np.arange(16, dtype=np.uint8))
    bad = bytearray(vr); bad[-1] = 2                      # invalid bool
    with pytest.raises(KgvError):
        decode_rows(kr, ko, bytes(bad), vo)
    bad = bytearray(vr); bad[10] = 9                      # script length disagrees with the row length
    with pytest.raises(KgvError):
        decode_rows(kr, ko, bytes(bad), vo)
    with pytest.raises(KgvError):
        decode_rows(kr[:20], np.array([0, 20], dtype=np.uint64), vr, vo)   # key shorter than txid + 1 byte
