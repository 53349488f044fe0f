"""DAG replay on the GPU: the caller side of the hot path (calculate_utxo_state /
verify_expected_utxo_state, consensus/src/pipeline/virtual_processor/utxo_validation.rs:110-228).

Two schedules with identical results:

  blockwise   for every (merged) block in order: validate_transactions_in_parallel(Full) against the UTXO
              table, then UtxoDiff::add_transaction for the accepted ones.  This is the reference's order; a
              10-BPS block carries <= ~300 signatures, far too few to fill a B200.

  windowed    signatures are context free given the spent output, so the scripts of a whole
              WINDOW of future blocks are checked in one large batch (flags = SCRIPTS_ONLY) with entries taken
              from the table or from outputs created inside the window; the in-order pass then only runs the
              UTXO-context rules (flags = SKIP_SCRIPT_CHECKS) block after block, entirely asynchronously:
              the accept mask is computed on the device and fed straight into kgv_utxo_apply_accepted.
              A transaction is accepted iff its context rules pass in order AND its scripts passed in the
              pre-check; context errors take precedence exactly as in
              validate_populated_transaction_and_get_fee (tx_validation_in_utxo_context.rs:34-61).
"""
import ctypes

import numpy as np

from .txbatch import ENTRY_DTYPE, TxBatch, build_batch
from .validator import (FLAGS_FULL, FLAGS_SCRIPTS_ONLY, FLAGS_SKIP_SCRIPT_CHECKS, RESULT_DTYPE, TX_NEEDS_HOST_VM, TX_OK, TX_SKIPPED_COINBASE, GpuUtxoSet,
                        TransactionValidator)
from .verifier import _KgvTxBatch


class DagReplayer:
    def __init__(self, ctx, params, capacity_slots=1 << 20):
        self.ctx = ctx
        self.tv = TransactionValidator(ctx, params)
        self.us = GpuUtxoSet(ctx, capacity_slots)

    def close(self):
        self.us.close()

    # ---------------------------------------------------------------------------------------- blockwise
    def replay_blockwise(self, blocks, multiset=None):
        """blocks: iterable of (txs, pov_daa_score). Returns the list of per-block RESULT arrays.
        multiset: optional MuHash that follows the UTXO set the way UtxoProcessingContext.multiset_hash does
        (utxo_validation.rs:120,144): the coinbase and every accepted transaction of each block are combined into it."""
        from .muhash import MuHash
        out = []
        for txs, pov in blocks:
            b = txs if isinstance(txs, TxBatch) else build_batch(txs)
            res = self.tv.validate_transactions_in_parallel(self.us, b, pov, FLAGS_FULL)
            if (res["status"] == TX_NEEDS_HOST_VM).any():
                self._host_vm_with_table(b, res)
            acc = ((res["status"] == TX_OK) | (res["status"] == TX_SKIPPED_COINBASE)).astype(np.uint8)
            if multiset is not None:  # before the spent entries are erased
                multiset.combine(MuHash.from_transactions(self.ctx, b, acc, pov, utxo_set=self.us))
            self.us.add_transactions(b, acc, pov)
            out.append(res)
        return out

    def _host_vm_with_table(self, b, res):
        """non-standard scripts: populate from the table on the host side, then the host engine decides"""
        keys = np.concatenate([b.inputs["prev_txid"], b.inputs["prev_index"].astype("<u4").view(np.uint8).reshape(-1, 4)], axis=1)
        found, ent, scr = self.us.get(keys, script_stride=256)
        pb = _with_entries(b, found, ent, scr)
        self.tv.check_scripts_host(pb, res)

    # ---------------------------------------------------------------------------------------- windowed
    def replay_windowed(self, blocks):
        """blocks: list of (txs, pov) forming ONE window. Returns per-block RESULT arrays (same values as blockwise)."""
        import torch
        blocks = list(blocks)
        all_txs, ranges = [], []
        for txs, pov in blocks:
            ranges.append((len(all_txs), len(all_txs) + len(txs), pov))
            all_txs.extend(txs)
        b = build_batch(all_txs)
        # --- pre-check of every script in the window
        ids = self.ctx.tx_ids(b)
        keys = np.concatenate([b.inputs["prev_txid"], b.inputs["prev_index"].astype("<u4").view(np.uint8).reshape(-1, 4)], axis=1)
        found, ent, scr = self.us.get(keys, script_stride=128)
        pb = _with_entries(b, found, ent, scr, window_ids=ids)
        pre = self.tv.validate_populated_transactions(pb, 0, FLAGS_SCRIPTS_ONLY, host_vm=True)
        # --- in-order pass, device resident, no host round trip per block.  The library kernels and the torch ops that
        # build the accept mask must be ordered: both run on one dedicated stream.
        dev = torch.device("cuda", self.ctx.device)
        stream = torch.cuda.Stream(device=dev)
        rebased = b.txs.copy()
        for a, e, _ in ranges:  # per-block relative first_input / first_output
            rebased["first_input"][a:e] -= b.txs["first_input"][a]
            rebased["first_output"][a:e] -= b.txs["first_output"][a]
        script_ok_h = ((pre["status"] == TX_OK) | (pre["status"] == TX_SKIPPED_COINBASE)).astype(np.uint8)
        self.ctx.use_stream(stream.cuda_stream)
        try:
            with torch.cuda.stream(stream):
                t_txs = torch.from_numpy(rebased.view(np.uint8).reshape(-1)).to(dev)
                t_in = torch.from_numpy(b.inputs.view(np.uint8).reshape(-1)).to(dev)
                t_out = torch.from_numpy(b.outputs.view(np.uint8).reshape(-1)).to(dev)
                t_bytes = torch.from_numpy(b.arena).to(dev)
                t_res = torch.zeros(len(all_txs) * 16, dtype=torch.uint8, device=dev)
                script_ok = torch.from_numpy(script_ok_h).to(dev)
                lib, h = self.ctx._lib, self.ctx._h
                keep = []
                for a, e, pov in ranges:
                    i0 = int(b.txs["first_input"][a])
                    i1 = int(b.txs["first_input"][e - 1] + b.txs["n_inputs"][e - 1])
                    o0 = int(b.txs["first_output"][a])
                    o1 = int(b.txs["first_output"][e - 1] + b.txs["n_outputs"][e - 1])
                    cb = _KgvTxBatch(t_txs.data_ptr() + 72 * a, e - a, t_in.data_ptr() + 56 * i0, i1 - i0, t_out.data_ptr() + 24 * o0, o1 - o0, None,
                                     t_bytes.data_ptr(), len(b.arena))
                    rp = t_res.data_ptr() + 16 * a
                    self.ctx._check(lib.kgv_validate_txs(h, self.us._h, ctypes.byref(cb), int(pov), FLAGS_SKIP_SCRIPT_CHECKS, ctypes.byref(self.tv.params), rp))
                    st = t_res[16 * a:16 * e].view(-1, 16)[:, 12]
                    acc = (((st == TX_OK) & (script_ok[a:e] != 0)) | (st == TX_SKIPPED_COINBASE)).to(torch.uint8).contiguous()
                    keep.append(acc)  # keep alive until the stream has consumed it
                    self.ctx._check(lib.kgv_utxo_apply_accepted(h, self.us._h, ctypes.byref(cb), acc.data_ptr(), int(pov)))
                stream.synchronize()
                res_host = t_res.cpu().numpy().tobytes()
        finally:
            self.ctx.reset_stream()
        ctxres = np.frombuffer(res_host, dtype=RESULT_DTYPE).copy()
        # merge: context verdict first, else the script verdict of the pre-check
        final = ctxres.copy()
        use_pre = ctxres["status"] == TX_OK
        for f in ("status", "script_err", "fail_input"):
            final[f][use_pre] = pre[f][use_pre]
        return [final[a:e] for a, e, _ in ranges]


def _with_entries(b, found, ent, scr, window_ids=None):
    """Populated copy of batch `b`: entries from a table lookup (found/ent/scr) and, where missing, from outputs of
    transactions of the same window (window_ids = their tx ids).  Absent entries are flagged (pad_[0] = 1)."""
    n_in = len(b.inputs)
    stride = scr.shape[1] if n_in else 0
    base = len(b.arena)
    E = np.zeros(n_in, dtype=ENTRY_DTYPE)
    for f in ("amount", "block_daa_score", "script_len", "spk_version", "is_coinbase"):
        E[f] = ent[f]
    E["script_off"] = base + np.arange(n_in, dtype=np.uint32) * stride
    absent = found == 0
    if window_ids is not None and absent.any() and len(window_ids):
        # resolve prev_txid among the window's tx ids (vectorised: sort + searchsorted on 32-byte keys)
        ids_v = np.ascontiguousarray(window_ids).view("V32").reshape(-1)
        order = np.argsort(ids_v)
        sorted_ids = ids_v[order]
        want = np.ascontiguousarray(b.inputs["prev_txid"][absent]).view("V32").reshape(-1)
        pos = np.searchsorted(sorted_ids, want)
        pos[pos >= len(sorted_ids)] = 0
        hit = sorted_ids[pos] == want
        src_tx = order[pos]
        idx_abs = np.nonzero(absent)[0]
        pidx = b.inputs["prev_index"][absent]
        ok = hit & (pidx < b.txs["n_outputs"][src_tx])
        o = b.txs["first_output"][src_tx[ok]] + pidx[ok]
        sel = idx_abs[ok]
        E["amount"][sel] = b.outputs["value"][o]
        E["script_off"][sel] = b.outputs["script_off"][o]
        E["script_len"][sel] = b.outputs["script_len"][o]
        E["spk_version"][sel] = b.outputs["spk_version"][o]
        E["is_coinbase"][sel] = (b.txs["flags"][src_tx[ok]] & 1)
        absent = absent.copy()
        absent[sel] = False
    E["pad_"][:, 0] = absent.astype(np.uint8)
    arena = np.concatenate([b.arena, scr.reshape(-1), np.zeros(8, np.uint8)]) if n_in else b.arena
    return TxBatch(b.txs, b.inputs, b.outputs, E, np.ascontiguousarray(arena))
