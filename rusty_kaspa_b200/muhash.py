"""MuHash on the GPU: mirror of the reference's `kaspa_muhash::MuHash` (crypto/muhash/src/lib.rs:32-121) and of
`MuHashExtensions` (consensus/core/src/muhash.rs) over the C ABI (include/kgv.h, K8).

A `MuHash` object holds the two canonical residues (numerator, denominator) as 384-byte little-endian strings on the
host; every operation that multiplies runs on the device.  There is no CPU arithmetic here."""
import ctypes

import numpy as np

from .verifier import _c_batch

ELEMENT_BYTE_SIZE = 384
_ONE = (1).to_bytes(ELEMENT_BYTE_SIZE, "little")


def _buf(b=None):
    a = np.zeros(ELEMENT_BYTE_SIZE, dtype=np.uint8)
    if b is not None:
        a[:] = np.frombuffer(b, dtype=np.uint8)
    return a


class MuHash:
    def __init__(self, ctx, numerator=_ONE, denominator=_ONE):
        self.ctx = ctx
        self._lib = ctx._lib
        self.numerator = bytes(numerator)
        self.denominator = bytes(denominator)

    # -- lib.rs:61-74 / :77-88: any number of elements per call (one product tree per field)
    def update(self, add=(), remove=()):
        """add_element for every byte string in `add`, remove_element for every one in `remove`."""
        items = [bytes(x) for x in add] + [bytes(x) for x in remove]
        if not items:
            return self
        flags = np.array([0] * len(add) + [1] * len(remove), dtype=np.uint8)
        offs = np.zeros(len(items) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(x) for x in items])
        data = np.frombuffer(b"".join(items) + b"\x00" * 8, dtype=np.uint8)
        num, den = _buf(), _buf()
        self.ctx._check(self._lib.kgv_muhash_elements(self.ctx._h, data.ctypes.data, offs.ctypes.data, flags.ctypes.data, len(items), num.ctypes.data, den.ctypes.data))
        return self.combine(MuHash(self.ctx, num.tobytes(), den.tobytes()))

    def add_element(self, data):
        return self.update(add=[data])

    def remove_element(self, data):
        return self.update(remove=[data])

    # -- lib.rs:91-96
    def combine(self, other):
        na, da, nb, db = _buf(self.numerator), _buf(self.denominator), _buf(other.numerator), _buf(other.denominator)
        self.ctx._check(self._lib.kgv_muhash_combine(self.ctx._h, na.ctypes.data, da.ctypes.data, nb.ctypes.data, db.ctypes.data))
        self.numerator, self.denominator = na.tobytes(), da.tobytes()
        return self

    # -- lib.rs:98-115
    def _finalize(self):
        n, d, ser, h = _buf(self.numerator), _buf(self.denominator), _buf(), np.zeros(32, dtype=np.uint8)
        self.ctx._check(self._lib.kgv_muhash_finalize(self.ctx._h, n.ctypes.data, d.ctypes.data, ser.ctypes.data, h.ctypes.data))
        self.numerator, self.denominator = ser.tobytes(), _ONE  # normalize()
        return ser.tobytes(), h.tobytes()

    def serialize(self):
        return self._finalize()[0]

    def finalize(self):
        return self._finalize()[1]

    # -- consensus/core/src/muhash.rs:16-27 over a whole batch (utxo_validation.rs:282-309)
    @classmethod
    def from_transactions(cls, ctx, batch, accept, pov_daa_score, utxo_set=None):
        """MuHash::from_transaction of every tx with accept[i] != 0, combined.  Entries: batch.entries, or looked up in
        `utxo_set` (a GpuUtxoSet) when given."""
        acc = np.ascontiguousarray(accept, dtype=np.uint8)
        cb = _c_batch(batch, with_entries=utxo_set is None)
        num, den = _buf(), _buf()
        ctx._check(ctx._lib.kgv_muhash_txs(ctx._h, utxo_set._h if utxo_set is not None else None, ctypes.byref(cb), acc.ctypes.data, int(pov_daa_score),
                                           num.ctypes.data, den.ctypes.data))
        return cls(ctx, num.tobytes(), den.tobytes())

    @classmethod
    def of_utxo_set(cls, ctx, utxo_set):
        """MuHash::add_utxo over every live entry of the GPU UTXO table."""
        num = _buf()
        ctx._check(ctx._lib.kgv_utxo_muhash(ctx._h, utxo_set._h, num.ctypes.data))
        return cls(ctx, num.tobytes(), _ONE)


# ---- batched forms (a replay window's worth of chain blocks at once) ----
def finalize_batch(ctx, values768, want_serialized=False):
    """values768: (n, 768) uint8 array of (numerator || denominator) records -> (n, 32) array of MuHash::finalize() hashes
    [and the (n, 384) serialized values].  One modular inversion for the whole batch (kgv_muhash_finalize_batch)."""
    v = np.ascontiguousarray(values768, dtype=np.uint8).reshape(-1, 768)
    n = len(v)
    h = np.zeros((n, 32), dtype=np.uint8)
    ser = np.zeros((n, 384), dtype=np.uint8) if want_serialized else None
    if n:
        ctx._check(ctx._lib.kgv_muhash_finalize_batch(ctx._h, v.ctypes.data, v.ctypes.data + 384, n, 768, ser.ctypes.data if want_serialized else None, h.ctypes.data))
    return (h, ser) if want_serialized else h


def prefix_combine(ctx, values768, init=None):
    """running MuHash::combine chain: record i becomes init * record 0 * ... * record i (kgv_muhash_prefix_combine). init: a MuHash or None."""
    v = np.ascontiguousarray(values768, dtype=np.uint8).reshape(-1, 768).copy()
    ib = None
    if init is not None:
        ib = np.frombuffer(init.numerator + init.denominator, dtype=np.uint8).copy()
    if len(v):
        ctx._check(ctx._lib.kgv_muhash_prefix_combine(ctx._h, ib.ctypes.data if ib is not None else None, v.ctypes.data, len(v)))
    return v
