"""Pruning-point UTXO-set import, the data-parallel part of IBD's last step, over libkgv.

Reference: the syncee receives the pruning point's UTXO set in chunks; every chunk goes through
`Consensus::append_imported_pruning_point_utxos` (consensus/src/consensus/mod.rs:1070-1083: write_many into the pruning UTXO store +
MuHash::from_utxo over the chunk in parallel, combined into the running multiset), then
`VirtualStateProcessor::import_pruning_point_utxo_set` (consensus/src/pipeline/virtual_processor/processor.rs:1126-1200) finalizes the
multiset and compares it with the new pruning point's header.utxo_commitment (ImportedMultisetHashMismatch), copies the set into virtual's
UTXO set and validates the pruning point's own transactions against it (validate_transactions_in_parallel, Full flags, :1162-1172:
NewPruningPointTxErrors unless every non-coinbase transaction passes).

Here the GPU table IS the (pruning = virtual) UTXO set; the chunks arrive as the store's rows (store_rows.decode_rows) or as arrays.
"""
import ctypes

import numpy as np

from . import store_rows
from .muhash import MuHash
from .txbatch import ENTRY_DTYPE
from .validator import TransactionValidator, FLAGS_FULL


class ImportedMultisetHashMismatch(Exception):
    """processor.rs:1134-1139"""


class NewPruningPointTxErrors(Exception):
    """processor.rs:1169-1172"""


class PruningPointImport:
    def __init__(self, ctx, utxo_set, params=None):
        self.ctx, self.us = ctx, utxo_set
        self.tv = TransactionValidator(ctx, params)
        self.numerator = bytearray(384)
        self.numerator[0] = 1
        self.n_imported = 0

    def append_imported_pruning_point_utxos(self, keys36, entries, arena):
        """one chunk: entries written into the GPU table, their MuHash elements multiplied into the running multiset (kgv_utxo_import_chunk)"""
        keys36 = np.ascontiguousarray(keys36, dtype=np.uint8).reshape(-1, 36)
        entries = np.ascontiguousarray(entries, dtype=ENTRY_DTYPE)
        arena = np.ascontiguousarray(arena, dtype=np.uint8) if len(arena) else np.zeros(8, np.uint8)
        if not len(keys36):
            return
        buf = (ctypes.c_uint8 * 384).from_buffer(self.numerator)
        self.ctx._check(self.ctx._lib.kgv_utxo_import_chunk(self.ctx._h, self.us._h, keys36.ctypes.data, entries.ctypes.data, arena.ctypes.data, len(arena), len(keys36),
                                                            ctypes.addressof(buf)))
        self.n_imported += len(keys36)

    def append_rows(self, key_rows, key_off, value_rows, value_off):
        """the same for a chunk of RocksDB rows (UtxoKey / bincode UtxoEntry, store_rows.py)"""
        self.append_imported_pruning_point_utxos(*store_rows.decode_rows(key_rows, key_off, value_rows, value_off))

    def multiset(self):
        one = bytes([1]) + bytes(383)
        return MuHash(self.ctx, bytes(self.numerator), one)

    def import_pruning_point_utxo_set(self, utxo_commitment, pruning_point_batch, daa_score):
        """finalize + compare with the header's commitment, then validate the pruning point's own transactions against the imported set.
        Returns the per-transaction results of that validation."""
        got = self.multiset().finalize()
        if got != bytes(utxo_commitment):
            raise ImportedMultisetHashMismatch(f"expected {bytes(utxo_commitment).hex()}, imported {got.hex()}")
        res = self.tv.validate_transactions_in_parallel(self.us, pruning_point_batch, int(daa_score), FLAGS_FULL)
        ok = int(((res["status"] == 0)).sum())
        if ok < len(res) - 1:  # every non-coinbase transaction must pass (position 0 is the coinbase: skipped)
            raise NewPruningPointTxErrors(f"{len(res) - 1 - ok} transactions of the new pruning point fail against the imported UTXO set")
        return res
