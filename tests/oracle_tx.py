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
