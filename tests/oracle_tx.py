"""ctypes glue handing a rusty_kaspa_b200.txbatch.TxBatch to the C oracle (test infrastructure)."""
import ctypes

import numpy as np


class OkBatch(ctypes.Structure):
    _fields_ = [("txs", ctypes.c_void_p), ("n_txs", ctypes.c_size_t), ("inputs", ctypes.c_void_p), ("n_inputs", ctypes.c_size_t),
                ("outputs", ctypes.c_void_p), ("n_outputs", ctypes.c_size_t), ("bytes", ctypes.c_void_p), ("n_bytes", ctypes.c_size_t)]


def ok_batch(b):
    ob = OkBatch(b.txs.ctypes.data, len(b.txs), b.inputs.ctypes.data, len(b.inputs), b.outputs.ctypes.data, len(b.outputs),
                 b.arena.ctypes.data, len(b.arena))
    ob._keep = b
    return ob


def tx_ids(lib, b, threads=1):
    out = np.zeros((len(b.txs), 32), dtype=np.uint8)
    ob = ok_batch(b)
    lib.ok_tx_ids(ctypes.byref(ob), out.ctypes.data_as(ctypes.c_void_p), threads)
    return out


def tx_hashes(lib, b, threads=1):
    out = np.zeros((len(b.txs), 32), dtype=np.uint8)
    ob = ok_batch(b)
    lib.ok_tx_hashes(ctypes.byref(ob), out.ctypes.data_as(ctypes.c_void_p), threads)
    return out


def sighash(lib, b, tx, input_index, hash_type, ecdsa=False):
    out = ctypes.create_string_buffer(32)
    ob = ok_batch(b)
    lib.ok_sighash(ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(tx), ctypes.c_uint32(input_index),
                   ctypes.c_uint8(hash_type), int(ecdsa), out)
    return out.raw


class OkParams(ctypes.Structure):
    _fields_ = [("coinbase_maturity", ctypes.c_uint64), ("storage_mass_parameter", ctypes.c_uint64), ("max_sompi", ctypes.c_uint64)]


RESULT_DTYPE = np.dtype([("fee", "<u8"), ("fail_input", "<u4"), ("status", "u1"), ("script_err", "u1"), ("pad_", "u1", (2,))])
MAX_SOMPI = 29_000_000_000 * 100_000_000
TX_STATUS = {0: "Ok", 1: "MissingTxOutpoints", 2: "ImmatureCoinbaseSpend", 3: "InputAmountOverflow", 4: "InputAmountTooHigh", 5: "SpendTooHigh",
             6: "MassIncomputable", 7: "WrongMass", 8: "SequenceLockConditionsAreNotMet", 9: "SignatureInvalid", 10: "SignatureEmpty",
             11: "NeedsHostVm", 12: "Coinbase"}
SCRIPT_ERR = {0: "Ok", 1: "EvalFalse", 2: "NullFail", 3: "InvalidSignature", 4: "SigLength", 5: "PubKeyFormat", 6: "InvalidSigHashType",
              7: "ExceededSigOpLimit", 255: "NonStandard"}


def params(coinbase_maturity=100, storage_mass_parameter=10**12, max_sompi=MAX_SOMPI):
    return OkParams(coinbase_maturity, storage_mass_parameter, max_sompi)


def check_script_std(lib, b, tx, input_index):
    ob = ok_batch(b)
    return lib.ok_check_script_std(ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(tx), ctypes.c_uint32(input_index))


def validate_populated(lib, b, tx, pov, flags, p):
    ob = ok_batch(b)
    out = np.zeros(1, dtype=RESULT_DTYPE)
    lib.ok_validate_populated(ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(tx), ctypes.c_uint64(pov), int(flags),
                              ctypes.byref(p), out.ctypes.data_as(ctypes.c_void_p))
    return out[0]


def storage_mass(lib, b, tx, storm):
    ob = ok_batch(b)
    m = ctypes.c_uint64()
    rc = lib.ok_storage_mass(ctypes.byref(ob), b.entries.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(tx), ctypes.c_uint64(storm), ctypes.byref(m))
    return None if rc else m.value


class State:
    def __init__(self, lib):
        self.lib = lib
        lib.ok_state_new.restype = ctypes.c_void_p
        lib.ok_state_count.restype = ctypes.c_uint64
        self.h = ctypes.c_void_p(lib.ok_state_new())

    def validate(self, b, pov, flags, p, threads=4):
        ob = ok_batch(b)
        out = np.zeros(len(b.txs), dtype=RESULT_DTYPE)
        self.lib.ok_state_validate(self.h, ctypes.byref(ob), ctypes.c_uint64(pov), int(flags), ctypes.byref(p), out.ctypes.data_as(ctypes.c_void_p), threads)
        return out

    def accept(self, b, accept, pov):
        ob = ok_batch(b)
        a = np.ascontiguousarray(accept, dtype=np.uint8)
        if not hasattr(self, "_log"):
            self._log = []
        self._log.append((b, a.copy(), pov))  # lets a test rebuild an equal state (the C oracle has no clone)
        return self.lib.ok_state_accept(self.h, ctypes.byref(ob), a.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(pov))

    def commit(self):
        self.lib.ok_state_commit(self.h)

    def count(self):
        return int(self.lib.ok_state_count(self.h))

    def digest(self):
        o = ctypes.create_string_buffer(32)
        self.lib.ok_state_digest(self.h, o)
        return o.raw

    def get(self, key36):
        e = np.zeros(1, dtype=np.dtype([("amount", "<u8"), ("block_daa_score", "<u8"), ("script_off", "<u4"), ("script_len", "<u4"),
                                        ("spk_version", "<u2"), ("is_coinbase", "u1"), ("pad_", "u1", (5,))]))
        sc = ctypes.create_string_buffer(10000)
        if not self.lib.ok_state_get(self.h, key36, e.ctypes.data_as(ctypes.c_void_p), sc, 10000):
            return None
        return {"amount": int(e[0]["amount"]), "block_daa_score": int(e[0]["block_daa_score"]), "spk_version": int(e[0]["spk_version"]),
                "is_coinbase": bool(e[0]["is_coinbase"]), "script": sc.raw[:int(e[0]["script_len"])]}

    def close(self):
        if self.h:
            self.lib.ok_state_free(self.h)
            self.h = None


def state_replay(state, b, block_first_tx, block_pov, p, block_flags=None, threads=4):
    """ok_state_replay: the whole window in one call. Returns (results, accept)."""
    ob = ok_batch(b)
    res = np.zeros(len(b.txs), dtype=RESULT_DTYPE)
    acc = np.zeros(len(b.txs), dtype=np.uint8)
    first = np.ascontiguousarray(block_first_tx, dtype=np.uint32)
    pov = np.ascontiguousarray(block_pov, dtype=np.uint64)
    fl = None if block_flags is None else np.ascontiguousarray(block_flags, dtype=np.uint32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = state.lib.ok_state_replay(state.h, ctypes.byref(ob), vp(first), vp(pov), None if fl is None else vp(fl), ctypes.c_size_t(len(pov)), ctypes.byref(p), vp(res), vp(acc),
                                   int(threads))
    assert rc == 0, "UtxoAlgebraError in the oracle replay"
    return res, acc
