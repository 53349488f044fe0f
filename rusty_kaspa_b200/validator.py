"""Host-side Python mirror of the validation surface of the reference, on top of the C ABI.

    TransactionValidator.validate_populated_transactions  <->  validate_populated_transaction_and_get_fee
        (consensus/src/processes/transaction_validator/tx_validation_in_utxo_context.rs:34-61)
    GpuUtxoSet (get / apply_diff / count / digest)          <->  UtxoView::get, DbUtxoSetStore::write_diff_batch
        (consensus/core/src/utxo/utxo_view.rs:5-7, consensus/src/model/stores/utxo_set.rs:107-112,143-152)
    TransactionValidator.validate_transactions_in_parallel  <->  VirtualStateProcessor::validate_transactions_in_parallel
        (consensus/src/pipeline/virtual_processor/utxo_validation.rs:262-278)
    GpuUtxoSet.add_transactions                              <->  UtxoDiff::add_transaction (utxo_diff.rs:233-247)
"""
import ctypes

import numpy as np

from . import _lib
from .txbatch import ENTRY_DTYPE
from .verifier import _c_batch

RESULT_DTYPE = np.dtype([("fee", "<u8"), ("fail_input", "<u4"), ("status", "u1"), ("script_err", "u1"), ("pad_", "u1", (2,))])
assert RESULT_DTYPE.itemsize == 16

FLAGS_FULL, FLAGS_SKIP_SCRIPT_CHECKS, FLAGS_SKIP_MASS_CHECK, FLAGS_SCRIPTS_ONLY = 0, 1, 2, 3
MAX_SOMPI = 29_000_000_000 * 100_000_000
TX_OK, TX_NEEDS_HOST_VM, TX_SKIPPED_COINBASE = 0, 11, 12


class Params(ctypes.Structure):
    """kgv_params: the consensus parameters the path reads (consensus/core/src/config/params.rs)."""
    _fields_ = [("coinbase_maturity", ctypes.c_uint64), ("storage_mass_parameter", ctypes.c_uint64), ("max_sompi", ctypes.c_uint64)]

    def __init__(self, coinbase_maturity=100, storage_mass_parameter=10**12, max_sompi=MAX_SOMPI):
        super().__init__(coinbase_maturity, storage_mass_parameter, max_sompi)


class SigRequest(ctypes.Structure):
    """kgv_sig_request: one signature check the host script engine asks a verdict for."""
    _fields_ = [("tx", ctypes.c_uint32), ("input", ctypes.c_uint32), ("hash_type", ctypes.c_uint8), ("ecdsa", ctypes.c_uint8), ("key_len", ctypes.c_uint8),
                ("pad_", ctypes.c_uint8), ("key", ctypes.c_uint8 * 33), ("sig", ctypes.c_uint8 * 64), ("pad2_", ctypes.c_uint8 * 3)]


assert ctypes.sizeof(SigRequest) == 112
VERDICT_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(SigRequest))

SCRIPT_ERR_NAMES = {0: "Ok", 1: "EvalFalse", 2: "NullFail", 3: "InvalidSignature", 4: "SigLength", 5: "PubKeyFormat", 6: "InvalidSigHashType",
                    7: "ExceededSigOpLimit", 8: "SignatureScriptNotPushOnly", 9: "CleanStack", 10: "EmptyStack", 11: "ElementTooBig",
                    12: "TooManyOperations", 13: "StackSizeExceeded", 14: "OpcodeDisabled", 15: "OpcodeReserved", 16: "InvalidOpcode",
                    17: "MalformedPush", 18: "MalformedPushSize", 19: "NotMinimalData", 20: "ErrUnbalancedConditional",
                    21: "InvalidState(condition stack empty)", 22: "InvalidState(expected boolean)", 23: "InvalidState(pick at an invalid location)",
                    24: "InvalidState(roll at an invalid location)", 25: "VerifyError", 26: "EarlyReturn", 27: "InvalidStackOperation", 28: "NumberTooBig",
                    29: "InvalidPubKeyCount", 30: "InvalidSignatureCount", 31: "UnsatisfiedLockTime", 32: "ScriptSize", 33: "NoScripts",
                    34: "InvalidInputIndex", 35: "InvalidOutputIndex", 36: "Serialization", 254: "NeedsSigVerdicts", 255: "NonStandard"}


def script_execute(batch, tx, input_index, verdict=None):
    """TxScriptEngine::from_transaction_input(..).execute() on the HOST engine of libkgv (no GPU involved).
    verdict(request: SigRequest) -> KGV_SIG_* or -1.  Returns the KGV_SCRIPT_* code."""
    lib = _lib.load()
    cb = _c_batch(batch, with_entries=True)
    fn = VERDICT_FN((lambda user, rq: int(verdict(rq.contents))) if verdict else (lambda user, rq: -1))
    err = ctypes.c_uint8(0)
    rc = lib.kgv_script_execute(ctypes.byref(cb), int(tx), int(input_index), ctypes.cast(fn, ctypes.c_void_p), None, ctypes.addressof(err))
    if rc != 0:
        raise _lib.KgvError(f"kgv_script_execute failed ({rc})")
    return int(err.value)


class GpuUtxoSet:
    """GPU-resident UTXO set (kgv_utxo_table)."""

    def __init__(self, ctx, capacity_slots=1 << 20):
        self.ctx = ctx
        self._lib = ctx._lib
        h = ctypes.c_void_p()
        ctx._check(self._lib.kgv_utxo_create(ctx._h, int(capacity_slots), ctypes.byref(h)))
        self._h = h

    def close(self):
        if self._h:
            self._lib.kgv_utxo_destroy(self.ctx._h, self._h)
            self._h = None

    # ---- composed views (utxo_view.rs:22-35): a diff layer over this set
    def compose(self, capacity_slots=1 << 16):
        """UtxoViewComposition::compose: a GpuUtxoSet that behaves as self ∘ (an initially empty diff); writes through it never touch self"""
        v = GpuUtxoSet.__new__(GpuUtxoSet)
        v.ctx, v._lib = self.ctx, self._lib
        h = ctypes.c_void_p()
        self.ctx._check(self._lib.kgv_utxo_view_create(self.ctx._h, self._h, int(capacity_slots), ctypes.byref(h)))
        v._h, v.base = h, self
        return v

    def commit(self):
        """fold this diff layer into the set below it (write_diff_batch) and empty it"""
        self.ctx._check(self._lib.kgv_utxo_view_commit(self.ctx._h, self._h))

    def discard(self):
        self.ctx._check(self._lib.kgv_utxo_view_discard(self.ctx._h, self._h))

    def get(self, keys36, script_stride=128):
        """keys36: (n, 36) uint8. Returns (found (n,), entries (n,) ENTRY_DTYPE, scripts (n, stride))."""
        keys36 = np.ascontiguousarray(keys36, dtype=np.uint8).reshape(-1, 36)
        n = len(keys36)
        ent = np.zeros(n, dtype=ENTRY_DTYPE)
        scr = np.zeros((n, script_stride), dtype=np.uint8)
        found = np.zeros(n, dtype=np.uint8)
        if n:
            self.ctx._check(self._lib.kgv_utxo_lookup(self.ctx._h, self._h, keys36.ctypes.data, n, ent.ctypes.data, scr.ctypes.data, script_stride, found.ctypes.data))
        return found, ent, scr

    def apply_diff(self, rem_keys36=None, add_keys36=None, add_entries=None, add_bytes=None):
        """write_diff_batch: delete `rem_keys36`, then put (add_keys36, add_entries[script_off/len into add_bytes])."""
        rk = np.zeros((0, 36), np.uint8) if rem_keys36 is None else np.ascontiguousarray(rem_keys36, dtype=np.uint8).reshape(-1, 36)
        ak = np.zeros((0, 36), np.uint8) if add_keys36 is None else np.ascontiguousarray(add_keys36, dtype=np.uint8).reshape(-1, 36)
        ae = np.zeros(0, ENTRY_DTYPE) if add_entries is None else np.ascontiguousarray(add_entries)
        ab = np.zeros(8, np.uint8) if add_bytes is None else np.ascontiguousarray(add_bytes, dtype=np.uint8)
        rs, as_ = np.zeros(max(len(rk), 1), np.uint8), np.zeros(max(len(ak), 1), np.uint8)
        self.ctx._check(self._lib.kgv_utxo_apply_diff(self.ctx._h, self._h, rk.ctypes.data if len(rk) else None, len(rk), rs.ctypes.data,
                                                      ak.ctypes.data if len(ak) else None, ae.ctypes.data if len(ak) else None, ab.ctypes.data, len(ab), len(ak),
                                                      as_.ctypes.data))
        return rs[:len(rk)], as_[:len(ak)]

    def add_transactions(self, batch, accept, pov_daa_score):
        """UtxoDiff::add_transaction for every tx with accept[i] != 0, applied to the table."""
        acc = np.ascontiguousarray(accept, dtype=np.uint8)
        cb = _c_batch(batch, with_entries=False)
        self.ctx._check(self._lib.kgv_utxo_apply_accepted(self.ctx._h, self._h, ctypes.byref(cb), acc.ctypes.data, int(pov_daa_score)))

    def count(self):
        c = ctypes.c_uint64()
        self.ctx._check(self._lib.kgv_utxo_count(self.ctx._h, self._h, ctypes.byref(c)))
        return int(c.value)

    def digest(self):
        out = (ctypes.c_uint8 * 32)()
        self.ctx._check(self._lib.kgv_utxo_digest(self.ctx._h, self._h, ctypes.addressof(out)))
        return bytes(out)

    def export(self):
        """DbUtxoSetStore::iterator (utxo_set.rs:114-129): every live entry. Returns (keys36 (n, 36), entries (n,) ENTRY_DTYPE, arena bytes)."""
        n, nb = ctypes.c_size_t(), ctypes.c_size_t()
        self.ctx._check(self._lib.kgv_utxo_export(self.ctx._h, self._h, None, None, None, 0, 0, ctypes.byref(n), ctypes.byref(nb)))
        keys = np.zeros((max(n.value, 1), 36), dtype=np.uint8)
        ent = np.zeros(max(n.value, 1), dtype=ENTRY_DTYPE)
        arena = np.zeros(max(nb.value, 8), dtype=np.uint8)
        self.ctx._check(self._lib.kgv_utxo_export(self.ctx._h, self._h, keys.ctypes.data, ent.ctypes.data, arena.ctypes.data, n.value, nb.value, ctypes.byref(n), ctypes.byref(nb)))
        return keys[:n.value], ent[:n.value], arena[:nb.value]


class TransactionValidator:
    """Batch counterpart of the reference's TransactionValidator for the UTXO-context rules."""

    def __init__(self, ctx, params=None):
        self.ctx = ctx
        self._lib = ctx._lib
        self.params = params or Params()

    def validate_populated_transactions(self, batch, pov_daa_score, flags=FLAGS_FULL, host_vm=False):
        """batch.entries must hold the populated UtxoEntry of every input. Returns RESULT_DTYPE[n_txs].
        host_vm=True additionally sends every KGV_TX_NEEDS_HOST_VM transaction through the host script engine."""
        res = np.zeros(batch.n_txs, dtype=RESULT_DTYPE)
        cb = _c_batch(batch, with_entries=True)
        self.ctx._check(self._lib.kgv_validate_populated(self.ctx._h, ctypes.byref(cb), int(pov_daa_score), int(flags), ctypes.byref(self.params), res.ctypes.data))
        if host_vm:
            self.check_scripts_host(batch, res)
        return res

    def check_scripts_host(self, batch, results):
        """check_scripts with the full host engine (GPU-verified signatures) for every tx whose status is NEEDS_HOST_VM; updates `results` in place."""
        idx = np.nonzero(results["status"] == TX_NEEDS_HOST_VM)[0].astype(np.uint32)
        if len(idx) == 0:
            return results
        out = np.zeros(len(idx), dtype=RESULT_DTYPE)
        cb = _c_batch(batch, with_entries=True)
        self.ctx._check(self._lib.kgv_check_scripts_host(self.ctx._h, ctypes.byref(cb), idx.ctypes.data, len(idx), out.ctypes.data))
        fees = results["fee"][idx]
        results[idx] = out
        results["fee"][idx] = fees
        return results

    def validate_mempool_transactions_in_parallel(self, utxo_set, batch, virtual_daa_score, flags=FLAGS_FULL):
        """consensus/src/pipeline/virtual_processor/processor.rs:853-878: the same UTXO-context validation run for a batch of
        mempool transactions against the virtual UTXO set; unlike the block path the per-transaction outcome is RETURNED
        (Vec<TxResult<()>>), not filtered: status / script_err / fail_input say why, `fee` feeds the host-side feerate check
        (tx_validation_in_utxo_context.rs:63-73, f64, stays on the host).  Partially populated transactions (orphans) show up as
        MISSING_OUTPOINTS.  Below a few hundred signature checks per call the CPU path is faster."""
        return self.validate_transactions_in_parallel(utxo_set, batch, virtual_daa_score, flags)

    def validate_transactions_with_muhash_in_parallel(self, utxo_set, batch, pov_daa_score, flags=FLAGS_FULL):
        """utxo_validation.rs:282-309: as validate_transactions_in_parallel, plus the combined MuHash::from_transaction of the
        accepted transactions.  Returns (RESULT_DTYPE[n_txs], MuHash)."""
        from .muhash import MuHash
        res = self.validate_transactions_in_parallel(utxo_set, batch, pov_daa_score, flags)
        return res, MuHash.from_transactions(self.ctx, batch, (res["status"] == 0).astype(np.uint8), pov_daa_score, utxo_set=utxo_set)

    def validate_transactions_in_parallel(self, utxo_set, batch, pov_daa_score, flags=FLAGS_FULL):
        """Populate from the GPU UTXO set, then validate. Returns RESULT_DTYPE[n_txs] (coinbase: status 12)."""
        res = np.zeros(batch.n_txs, dtype=RESULT_DTYPE)
        cb = _c_batch(batch, with_entries=False)
        self.ctx._check(self._lib.kgv_validate_txs(self.ctx._h, utxo_set._h, ctypes.byref(cb), int(pov_daa_score), int(flags), ctypes.byref(self.params), res.ctypes.data))
        return res


class SigCache:
    """Device-resident verdict cache (kgv_sigcache): the counterpart of `Cache<SigCacheKey, bool>` (crypto/txscript/src/caches.rs:14-55).
    attach(ctx) makes every validation call of that context consult and fill it."""

    def __init__(self, ctx, capacity=10_000):
        self.ctx = ctx
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.kgv_sigcache_create(ctx._h, int(capacity), ctypes.byref(h)))
        self._h = h

    def attach(self, ctx=None):
        c = ctx or self.ctx
        c._check(c._lib.kgv_set_sigcache(c._h, self._h))

    def detach(self, ctx=None):
        c = ctx or self.ctx
        c._check(c._lib.kgv_set_sigcache(c._h, None))

    def clear(self):
        self.ctx._check(self.ctx._lib.kgv_sigcache_clear(self.ctx._h, self._h))

    def counters(self):
        v = [ctypes.c_uint64() for _ in range(4)]
        self.ctx._check(self.ctx._lib.kgv_sigcache_counters(self.ctx._h, self._h, *[ctypes.byref(x) for x in v]))
        return dict(zip(("hits", "inserts", "lookups", "evictions"), (int(x.value) for x in v)))

    def close(self):
        if self._h:
            self.detach()
            self.ctx._lib.kgv_sigcache_destroy(self._h)
            self._h = None
