"""RocksDB row formats of the reference's UTXO store over the C ABI: kgv_utxo_rows_encode / kgv_utxo_rows_decode.
  key row    consensus/src/model/stores/utxo_set.rs:31-62   txid || index LE, trailing zero index bytes trimmed (>= 1 kept)
  value row  bincode(UtxoEntry)  database/src/access.rs:139, consensus/core/src/tx.rs:49-57
Host-side functions (no GPU involved): they sit between the database and GpuUtxoSet.apply_diff / get."""
import ctypes

import numpy as np

from . import _lib
from .txbatch import ENTRY_DTYPE


def encode_rows(keys36, entries, arena):
    """-> (key_rows bytes, key_off (n+1,) u64, value_rows bytes, value_off (n+1,) u64)"""
    lib = _lib.load()
    keys36 = np.ascontiguousarray(keys36, dtype=np.uint8).reshape(-1, 36)
    entries = np.ascontiguousarray(entries, dtype=ENTRY_DTYPE)
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    n = len(keys36)
    ko, vo = np.zeros(n + 1, dtype=np.uint64), np.zeros(n + 1, dtype=np.uint64)
    rc = lib.kgv_utxo_rows_encode(keys36.ctypes.data, entries.ctypes.data, arena.ctypes.data, len(arena), n, None, ko.ctypes.data, None, vo.ctypes.data, 0, 0)
    if rc:
        raise _lib.KgvError(f"kgv_utxo_rows_encode failed ({rc})")
    kr, vr = np.zeros(int(ko[n]) + 1, dtype=np.uint8), np.zeros(int(vo[n]) + 1, dtype=np.uint8)
    rc = lib.kgv_utxo_rows_encode(keys36.ctypes.data, entries.ctypes.data, arena.ctypes.data, len(arena), n, kr.ctypes.data, ko.ctypes.data, vr.ctypes.data, vo.ctypes.data,
                                  int(ko[n]), int(vo[n]))
    if rc:
        raise _lib.KgvError(f"kgv_utxo_rows_encode failed ({rc})")
    return kr[:int(ko[n])].tobytes(), ko, vr[:int(vo[n])].tobytes(), vo


def decode_rows(key_rows, key_off, value_rows, value_off):
    """-> (keys36 (n,36), entries ENTRY_DTYPE[n], arena uint8[...])"""
    lib = _lib.load()
    n = len(key_off) - 1
    kr = np.frombuffer(bytes(key_rows) + b"\x00", dtype=np.uint8)
    vr = np.frombuffer(bytes(value_rows) + b"\x00", dtype=np.uint8)
    ko, vo = np.ascontiguousarray(key_off, dtype=np.uint64), np.ascontiguousarray(value_off, dtype=np.uint64)
    keys = np.zeros((n, 36), dtype=np.uint8)
    ent = np.zeros(n, dtype=ENTRY_DTYPE)
    arena = np.zeros(len(vr) + 8, dtype=np.uint8)
    used = ctypes.c_size_t()
    rc = lib.kgv_utxo_rows_decode(kr.ctypes.data, ko.ctypes.data, vr.ctypes.data, vo.ctypes.data, n, keys.ctypes.data, ent.ctypes.data, arena.ctypes.data, len(arena), ctypes.byref(used))
    if rc:
        raise _lib.KgvError(f"kgv_utxo_rows_decode failed ({rc}): malformed row")
    return keys, ent, arena[:int(used.value) + 8]
