"""DAG replay on the GPU: the caller side of the hot path (calculate_utxo_state /
verify_expected_utxo_state, consensus/src/pipeline/virtual_processor/utxo_validation.rs:110-228).

Two schedules with identical results:

  blockwise   for every (merged) block in order: validate_transactions_in_parallel(Full) against the UTXO
              table, then UtxoDiff::add_transaction for the accepted ones.  This is the reference's order; a
              10-BPS block carries <= ~300 signatures, far too few to fill a B200.

  windowed    ONE library call per window of blocks (kgv_replay_window, include/kgv.h): every script of the window
              is checked in one large batch (signatures are context free given the spent output; outputs
              created inside the window are resolved on the device), then a single persistent kernel walks the
              blocks in order: populate, UTXO-context rules, accept = context ok AND scripts ok, erase / insert.
              Nothing of the schedule lives in Python any more: this module only flattens the blocks into the
              batch layout and names the per-block flags.
"""
import ctypes

import numpy as np

from .txbatch import TxBatch, build_batch
from .validator import FLAGS_FULL, RESULT_DTYPE, TX_OK, TX_SKIPPED_COINBASE, GpuUtxoSet, TransactionValidator
from .verifier import _c_batch

REPLAY_BLOCK_DTYPE = np.dtype([("first_tx", "<u4"), ("n_txs", "<u4"), ("pov_daa_score", "<u8"), ("flags", "<u4"), ("pad_", "<u4")])
assert REPLAY_BLOCK_DTYPE.itemsize == 24
REPLAY_ACCEPT_COINBASE, REPLAY_SKIP_SCRIPTS, REPLAY_VERIFY_ONLY = 1, 2, 4


class ReplayStats(ctypes.Structure):
    _fields_ = [("n_accepted", ctypes.c_uint64), ("n_sig_checks", ctypes.c_uint64), ("n_host_vm", ctypes.c_uint64), ("pre_check_ms", ctypes.c_float),
                ("in_order_ms", ctypes.c_float)]


def replay_blocks_array(ranges):
    """ranges: iterable of (first_tx, n_txs, pov_daa_score, flags) -> REPLAY_BLOCK_DTYPE array"""
    ranges = list(ranges)
    a = np.zeros(len(ranges), dtype=REPLAY_BLOCK_DTYPE)
    for i, (f, n, pov, fl) in enumerate(ranges):
        a[i] = (f, n, pov, fl, 0)
    return a


class DagReplayer:
    """Replays a linearised schedule of blocks.  Every block is given as (txs, pov_daa_score[, flags]); txs[0] is the
    block's coinbase (skipped by position, utxo_validation.rs:273).  flags default to REPLAY_ACCEPT_COINBASE: the
    generated schedules are chains (one merged block = the selected parent per chain block, its coinbase accepted,
    utxo_validation.rs:116-121) that are nevertheless fully script-checked.  For a real DAG the caller marks only the
    selected parent of each mergeset with ACCEPT_COINBASE | SKIP_SCRIPTS and gives all merged blocks the chain block's
    pov_daa_score (tests/test_gpu_replay.py replays the reference's simpa fixtures that way)."""

    def __init__(self, ctx, params, capacity_slots=1 << 20):
        self.ctx = ctx
        self.tv = TransactionValidator(ctx, params)
        self.us = GpuUtxoSet(ctx, capacity_slots)
        self.last_stats = None

    def close(self):
        self.us.close()

    # ---------------------------------------------------------------------------------------- blockwise
    def replay_blockwise(self, blocks, multiset=None):
        """Returns the list of per-block RESULT arrays.
        multiset: optional MuHash that follows the UTXO set the way UtxoProcessingContext.multiset_hash does
        (utxo_validation.rs:120,144): the accepted coinbase and every accepted transaction of each block are combined into it."""
        from .muhash import MuHash
        out = []
        for blk in blocks:
            txs, pov = blk[0], blk[1]
            flags = blk[2] if len(blk) > 2 else REPLAY_ACCEPT_COINBASE
            b = txs if isinstance(txs, TxBatch) else build_batch(txs)
            vflags = 1 if (flags & REPLAY_SKIP_SCRIPTS) else FLAGS_FULL
            res = self.tv.validate_transactions_in_parallel(self.us, b, pov, vflags)  # non-standard scripts are decided inside the call
            res["status"][0] = TX_SKIPPED_COINBASE  # position 0 is the coinbase
            acc = (res["status"] == TX_OK).astype(np.uint8)
            acc[0] = 1 if (flags & REPLAY_ACCEPT_COINBASE) else 0
            if flags & REPLAY_VERIFY_ONLY:
                acc[:] = 0
            if multiset is not None and acc.any():  # before the spent entries are erased
                multiset.combine(MuHash.from_transactions(self.ctx, b, acc, pov, utxo_set=self.us))
            if acc.any():
                self.us.add_transactions(b, acc, pov)
            out.append(res)
        return out

    # ---------------------------------------------------------------------------------------- windowed
    def replay_window(self, batch, blocks_arr, want_accept=False):
        """One kgv_replay_window call on a flattened batch; returns RESULT_DTYPE[n_txs] (and the accept mask)."""
        res = np.zeros(batch.n_txs, dtype=RESULT_DTYPE)
        acc = np.zeros(batch.n_txs, dtype=np.uint8)
        st = ReplayStats()
        cb = _c_batch(batch, with_entries=False)
        blocks_arr = np.ascontiguousarray(blocks_arr, dtype=REPLAY_BLOCK_DTYPE)
        self.ctx._check(self.ctx._lib.kgv_replay_window(self.ctx._h, self.us._h, ctypes.byref(cb), blocks_arr.ctypes.data, len(blocks_arr),
                                                        ctypes.byref(self.tv.params), res.ctypes.data, acc.ctypes.data, ctypes.byref(st)))
        self.last_stats = {"n_accepted": int(st.n_accepted), "n_sig_checks": int(st.n_sig_checks), "n_host_vm": int(st.n_host_vm), "pre_check_ms": float(st.pre_check_ms),
                           "in_order_ms": float(st.in_order_ms)}
        return (res, acc) if want_accept else res

    def prefetch(self, batch):
        """kgv_batch_prefetch: start uploading the NEXT window's (host) batch while the current window computes; the replay_window call that gets
        this very batch object then skips its upload"""
        cb = _c_batch(batch, with_entries=False)
        self.ctx._check(self.ctx._lib.kgv_batch_prefetch(self.ctx._h, ctypes.byref(cb)))

    def replay_muhash(self, group_first_block):
        """kgv_replay_muhash for the window just replayed: (n_groups, 768) uint8 (numerator || denominator) of what each group of blocks accepted"""
        gf = np.ascontiguousarray(group_first_block, dtype=np.uint32)
        out = np.zeros((len(gf) - 1, 768), dtype=np.uint8)
        if len(gf) > 1:
            self.ctx._check(self.ctx._lib.kgv_replay_muhash(self.ctx._h, gf.ctypes.data, len(gf) - 1, out.ctypes.data))
        return out

    def replay_windowed(self, blocks):
        """blocks: list of (txs, pov[, flags]) forming ONE window. Returns per-block RESULT arrays (same values as blockwise)."""
        blocks = list(blocks)
        all_txs, ranges = [], []
        for blk in blocks:
            txs, pov = blk[0], blk[1]
            flags = blk[2] if len(blk) > 2 else REPLAY_ACCEPT_COINBASE
            ranges.append((len(all_txs), len(txs), pov, flags))
            all_txs.extend(txs)
        b = build_batch(all_txs)
        res = self.replay_window(b, replay_blocks_array(ranges))
        return [res[f:f + n] for f, n, _, _ in ranges]
