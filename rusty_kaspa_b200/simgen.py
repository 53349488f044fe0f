"""simpa-shaped DAG workload generator (BASELINE configs 1, 3, 4, 5), seeded and self-contained.

Restates the *shape* of what simpa's miner emits (simpa/src/simulator/miner.rs:138-207: native-subnetwork
version-0 transactions spending the miner's own earlier outputs, P2PK Schnorr `20 <xonly> ac` outputs,
`41 <sig64> 01` signature scripts, sig_op_count = 1, storage mass committed with C from
simpa/src/main.rs:205, <= --tpb transactions per block) and extends it as BASELINE.json asks:
2-in/2-out transactions (config 3), P2PK-ECDSA and P2SH 2-of-3 multisig (config 4), and a small
fraction of deliberately invalid transactions.  simpa itself uses thread_rng, so its exact DAGs are
not reproducible; this generator is.

Blocks are produced as a linearised schedule (one merged block per "chain block", pov_daa_score =
block index): GHOSTDAG ordering is out of scope, the consumer replays the schedule.

No elliptic-curve library and nothing from oracle/ is used: keys and nonces come from
workload.ScalarPointPool, signatures are scalar arithmetic over those pools.
"""
import hashlib
import struct

import numpy as np

from . import workload as W

N = W.N
SUBNET_NATIVE = bytes(20)
SUBNET_COINBASE = bytes([1]) + bytes(19)
SIGHASH_ALL = 1
DEFAULT_STORAGE_MASS_PARAMETER = 10_000  # simpa/src/main.rs:205
DEFAULT_COINBASE_MATURITY = 200          # simpa/src/main.rs:204

KIND_P2PK, KIND_P2PK_ECDSA, KIND_MS, KIND_MS_ECDSA = 0, 1, 2, 3


def _h(domain, data):
    return hashlib.blake2b(data, digest_size=32, key=domain).digest()


def _varbytes(b):
    return struct.pack("<Q", len(b)) + b


def _enc_output(o):
    return struct.pack("<QH", o["value"], o["spk_version"]) + _varbytes(o["script"])


def tx_id(tx):
    """hashing/tx.rs:30-42"""
    cb = tx["subnetwork_id"] == SUBNET_COINBASE
    b = struct.pack("<HQ", tx["version"], len(tx["inputs"]))
    for i in tx["inputs"]:
        b += i["txid"] + struct.pack("<I", i["index"])
        b += (_varbytes(i["sigscript"]) + bytes([i["sig_op_count"]])) if cb else _varbytes(b"")
        b += struct.pack("<Q", i["sequence"])
    b += struct.pack("<Q", len(tx["outputs"])) + b"".join(_enc_output(o) for o in tx["outputs"])
    b += struct.pack("<Q", tx["lock_time"]) + tx["subnetwork_id"] + struct.pack("<Q", tx["gas"]) + _varbytes(tx["payload"])
    if cb and tx.get("mass", 0) > 0:
        b += struct.pack("<Q", tx["mass"])
    return _h(b"TransactionID", b)


def sighash_all(tx, entries, idx, ecdsa):
    """hashing/sighash.rs:238-277 for SIG_HASH_ALL"""
    H = lambda d: _h(b"TransactionSigningHash", d)
    prev = H(b"".join(i["txid"] + struct.pack("<I", i["index"]) for i in tx["inputs"]))
    seqs = H(b"".join(struct.pack("<Q", i["sequence"]) for i in tx["inputs"]))
    sops = H(bytes(i["sig_op_count"] for i in tx["inputs"]))
    outs = H(b"".join(_enc_output(o) for o in tx["outputs"]))
    pay = bytes(32) if (tx["subnetwork_id"] == SUBNET_NATIVE and not tx["payload"]) else H(_varbytes(tx["payload"]))
    i, e = tx["inputs"][idx], entries[idx]
    pre = (struct.pack("<H", tx["version"]) + prev + seqs + sops + i["txid"] + struct.pack("<I", i["index"])
           + struct.pack("<H", e["spk_version"]) + _varbytes(e["script"]) + struct.pack("<QQ", e["amount"], i["sequence"])
           + bytes([i["sig_op_count"]]) + outs + struct.pack("<Q", tx["lock_time"]) + tx["subnetwork_id"]
           + struct.pack("<Q", tx["gas"]) + pay + bytes([SIGHASH_ALL]))
    d = H(pre)
    if ecdsa:
        d = hashlib.sha256(hashlib.sha256(b"TransactionSigningHashECDSA").digest() + d).digest()
    return d


def storage_mass(in_amounts_scriptlens, out_values_scriptlens, C):
    """consensus/core/src/mass/mod.rs:338-410 (non-coinbase, no overflow expected for generator values)"""
    plur = lambda l: (63 + l + 99) // 100
    outs_plur = sum(plur(l) for _, l in out_values_scriptlens)
    harm_outs = sum(C * plur(l) * plur(l) // v for v, l in out_values_scriptlens)
    ins_plur = sum(plur(l) for _, l in in_amounts_scriptlens)
    n_in = len(in_amounts_scriptlens)
    if outs_plur == 1 or (n_in <= 2 and (ins_plur == 1 or (outs_plur == 2 and ins_plur == 2))):
        harm_ins = sum(C * plur(l) * plur(l) // a for a, l in in_amounts_scriptlens)
        return max(0, harm_outs - harm_ins)
    mean = sum(a for a, _ in in_amounts_scriptlens) // ins_plur
    return max(0, harm_outs - ins_plur * (C // mean))


def _push(data):
    n = len(data)
    if n <= 75:
        return bytes([n]) + data
    if n <= 255:
        return bytes([0x4C, n]) + data
    return bytes([0x4D, n & 0xFF, n >> 8]) + data


class SimDag:
    """Seeded generator of blocks of signed transactions with a private view of the spendable outputs."""

    def __init__(self, seed=0x6B61737061, n_keys=1024, n_nonces=4096, storage_mass_parameter=DEFAULT_STORAGE_MASS_PARAMETER,
                 coinbase_maturity=DEFAULT_COINBASE_MATURITY, mix=(1.0, 0.0, 0.0, 0.0), frac_two_inputs=0.5, frac_invalid=0.0,
                 coinbase_outputs=8, subsidy=50_000_000_000):
        self.rng = np.random.default_rng(seed)
        self.keys = W.ScalarPointPool(n_keys, seed, b"sim-keys")
        self.nonces = W.ScalarPointPool(n_nonces, seed, b"sim-nonces")
        self.kinv = [pow(k, -1, N) for k in self.nonces.scalars]
        self.C = storage_mass_parameter
        self.maturity = coinbase_maturity
        self.mix = np.array(mix, dtype=float) / sum(mix)
        self.frac_two_inputs = frac_two_inputs
        self.frac_invalid = frac_invalid
        self.coinbase_outputs = coinbase_outputs
        self.subsidy = subsidy
        self.utxos = []  # dicts: txid, index, amount, script, kind, keys(list of key idx), redeem, daa, coinbase
        self.daa = 0
        self.n_signatures = 0

    # ---- scripts -------------------------------------------------------------------------------
    def _new_output_script(self):
        kind = int(self.rng.choice(4, p=self.mix))
        if kind == KIND_P2PK:
            k = int(self.rng.integers(0, self.keys.count))
            return kind, [k], None, bytes([0x20]) + self.keys.xs[k] + bytes([0xAC])
        if kind == KIND_P2PK_ECDSA:
            k = int(self.rng.integers(0, self.keys.count))
            return kind, [k], None, bytes([0x21, 0x02]) + self.keys.xs[k] + bytes([0xAB])
        ks = [int(x) for x in self.rng.choice(self.keys.count, size=3, replace=False)]
        if kind == KIND_MS:
            redeem = bytes([0x52]) + b"".join(bytes([0x20]) + self.keys.xs[k] for k in ks) + bytes([0x53, 0xAE])
        else:
            redeem = bytes([0x52]) + b"".join(bytes([0x21, 0x02]) + self.keys.xs[k] for k in ks) + bytes([0x53, 0xA9])
        spk = bytes([0xAA, 0x20]) + hashlib.blake2b(redeem, digest_size=32).digest() + bytes([0x87])
        return kind, ks, redeem, spk

    def _sign(self, key_idx, msg, ecdsa):
        j = int(self.rng.integers(0, self.nonces.count))
        d, k = self.keys.scalars[key_idx], self.nonces.scalars[j]
        self.n_signatures += 1
        if not ecdsa:
            e = W._challenge(self.nonces.xs[j], self.keys.xs[key_idx], msg)
            return self.nonces.xs[j] + ((k + e * d) % N).to_bytes(32, "big")
        r = int.from_bytes(self.nonces.xs[j], "big") % N
        s = self.kinv[j] * (int.from_bytes(msg, "big") + r * d) % N
        if s > N // 2:
            s = N - s
        return r.to_bytes(32, "big") + s.to_bytes(32, "big")

    # ---- blocks --------------------------------------------------------------------------------
    def make_block(self, n_txs):
        """Returns (txs, pov_daa_score): txs[0] is the coinbase; the rest spend earlier outputs."""
        self.daa += 1
        pov = self.daa
        txs = []
        cb_outs = []
        created = []
        for _ in range(self.coinbase_outputs):
            kind, ks, redeem, spk = self._new_output_script()
            cb_outs.append({"value": self.subsidy // self.coinbase_outputs, "spk_version": 0, "script": spk})
            created.append((kind, ks, redeem))
        cb = {"version": 0, "inputs": [], "outputs": cb_outs, "lock_time": 0, "subnetwork_id": SUBNET_COINBASE, "gas": 0,
              "payload": struct.pack("<Q", pov) + b"kgv-sim", "mass": 0}
        txs.append(cb)
        new_utxos = [self._utxo(cb, i, created[i], pov, True) for i in range(len(cb_outs))]
        spent_in_block = set()
        for _ in range(n_txs):
            tx = self._make_tx(pov, spent_in_block, new_utxos)
            if tx is not None:
                txs.append(tx)
        self.utxos.extend(new_utxos)
        return txs, pov

    def _utxo(self, tx, index, created, daa, coinbase, txid=None):
        kind, ks, redeem = created
        o = tx["outputs"][index]
        return {"txid": txid if txid is not None else tx_id(tx), "index": index, "amount": o["value"], "script": o["script"], "kind": kind,
                "keys": ks, "redeem": redeem, "daa": daa, "coinbase": coinbase}

    def _pick(self, pov, spent_in_block):
        for _ in range(50):
            if not self.utxos:
                return None
            i = int(self.rng.integers(0, len(self.utxos)))
            u = self.utxos[i]
            if u["coinbase"] and u["daa"] + self.maturity > pov:
                continue
            if u["amount"] < 4 or (u["txid"], u["index"]) in spent_in_block:
                continue
            self.utxos[i] = self.utxos[-1]
            self.utxos.pop()
            spent_in_block.add((u["txid"], u["index"]))
            return u
        return None

    def _make_tx(self, pov, spent_in_block, new_utxos):
        n_in = 2 if self.rng.random() < self.frac_two_inputs else 1
        ins = [u for u in (self._pick(pov, spent_in_block) for _ in range(n_in)) if u is not None]
        if not ins:
            return None
        total = sum(u["amount"] for u in ins)
        fee = 1
        created, outs = [], []
        v0 = (total - fee) // 2
        for v in (v0, total - fee - v0):
            kind, ks, redeem, spk = self._new_output_script()
            outs.append({"value": v, "spk_version": 0, "script": spk})
            created.append((kind, ks, redeem))
        tx = {"version": 0, "inputs": [{"txid": u["txid"], "index": u["index"], "sigscript": b"", "sequence": 0,
                                          "sig_op_count": 1 if u["kind"] in (KIND_P2PK, KIND_P2PK_ECDSA) else 3} for u in ins],
              "outputs": outs, "lock_time": 0, "subnetwork_id": SUBNET_NATIVE, "gas": 0, "payload": b"", "mass": 0}
        tx["mass"] = storage_mass([(u["amount"], len(u["script"])) for u in ins], [(o["value"], len(o["script"])) for o in outs], self.C)
        invalid = self.rng.random() < self.frac_invalid
        mode = int(self.rng.integers(0, 8)) if invalid else -1
        k0 = ins[0]["kind"]
        if (mode == 5 and k0 not in (KIND_P2PK_ECDSA,)) or (mode == 7 and k0 not in (KIND_MS, KIND_MS_ECDSA)) or (mode == 6 and k0 in (KIND_MS, KIND_MS_ECDSA)):
            mode = 4
        if mode == 0:
            tx["mass"] += 1                                   # WrongMass
        elif mode == 1:
            tx["outputs"][0]["value"] += total                 # SpendTooHigh
        elif mode == 2:
            tx["inputs"][0]["txid"] = bytes(self.rng.integers(0, 256, 32, dtype=np.uint8))  # MissingTxOutpoints
        elif mode == 3:
            tx["inputs"][0]["sig_op_count"] = 0                # ExceededSigOpLimit
        entries = [{"amount": u["amount"], "spk_version": 0, "script": u["script"]} for u in ins]
        for idx, u in enumerate(ins):
            ecdsa = u["kind"] in (KIND_P2PK_ECDSA, KIND_MS_ECDSA)
            msg = sighash_all(tx, entries, idx, ecdsa)
            if u["kind"] in (KIND_P2PK, KIND_P2PK_ECDSA):
                sig = self._sign(u["keys"][0], msg, ecdsa)
                if mode == 4 and idx == 0:
                    sig = sig[:40] + bytes([sig[40] ^ 1]) + sig[41:]   # EvalFalse
                if mode == 5 and idx == 0 and ecdsa:
                    sig = sig[:32] + (N - int.from_bytes(sig[32:], "big")).to_bytes(32, "big")  # high S
                ht = 0x03 if (mode == 6 and idx == 0) else SIGHASH_ALL  # InvalidSigHashType
                tx["inputs"][idx]["sigscript"] = bytes([0x41]) + sig + bytes([ht])
            else:
                pair = sorted(int(x) for x in self.rng.choice(3, size=2, replace=False))
                if mode == 7 and idx == 0:
                    pair = pair[::-1]                                   # wrong order => NullFail
                sigs = [self._sign(u["keys"][p], msg, ecdsa) for p in pair]
                if mode == 4 and idx == 0:
                    sigs[1] = sigs[1][:40] + bytes([sigs[1][40] ^ 1]) + sigs[1][41:]
                tx["inputs"][idx]["sigscript"] = b"".join(bytes([0x41]) + s + bytes([SIGHASH_ALL]) for s in sigs) + _push(u["redeem"])
        if mode == -1:  # outputs only become spendable for the generator if the tx is valid
            tid = tx_id(tx)
            for i in range(len(outs)):
                new_utxos.append(self._utxo(tx, i, created[i], pov, False, txid=tid))
        else:
            for u in ins:  # an invalid tx is not accepted: its inputs stay unspent
                self.utxos.append(u)
        return tx


def funded_window(n_txs, seed=0x6B61737061, n_keys=4096, n_nonces=4096, storage_mass_parameter=DEFAULT_STORAGE_MASS_PARAMETER, two_input_fraction=0.5,
                  mix=(1.0, 0.0, 0.0, 0.0)):
    """One pre-verification window of `n_txs` mutually independent transactions spending distinct funding outputs
    (what ~n_txs/150 consecutive 10-BPS blocks carry).  mix = fractions of (P2PK Schnorr, P2PK ECDSA, P2SH 2-of-3
    Schnorr multisig, P2SH 2-of-3 ECDSA multisig) among the SPENT outputs: (1,0,0,0) is BASELINE config 3,
    (0.5,0,0.25,0.25)-like mixes are config 4.  Returns (funding_keys36 (m,36) u8, funding entry dicts, tx dicts);
    the funding outputs are to be loaded into the UTXO set first (kgv_utxo_apply_diff)."""
    dag = SimDag(seed=seed, n_keys=n_keys, n_nonces=n_nonces, storage_mass_parameter=storage_mass_parameter, mix=mix)
    rng = dag.rng
    fund_keys, fund_entries, txs = [], [], []
    for t in range(n_txs):
        n_in = 2 if rng.random() < two_input_fraction else 1
        ins, ents, meta = [], [], []
        for _ in range(n_in):
            kind, ks, redeem, spk = dag._new_output_script()
            txid = hashlib.blake2b(struct.pack("<QQ", seed & 0xFFFFFFFFFFFF, len(fund_keys)), digest_size=32).digest()
            amount = int(rng.integers(10**8, 10**11))
            fund_keys.append(txid + struct.pack("<I", 0))
            fund_entries.append({"amount": amount, "spk_version": 0, "script": spk, "block_daa_score": 1, "is_coinbase": False})
            ins.append({"txid": txid, "index": 0, "sigscript": b"", "sequence": 0, "sig_op_count": 1 if kind in (KIND_P2PK, KIND_P2PK_ECDSA) else 3})
            ents.append(fund_entries[-1])
            meta.append((kind, ks, redeem))
        total = sum(e["amount"] for e in ents)
        outs = []
        for v in ((total - 1) // 2, total - 1 - (total - 1) // 2):
            k = int(rng.integers(0, dag.keys.count))
            outs.append({"value": v, "spk_version": 0, "script": bytes([0x20]) + dag.keys.xs[k] + bytes([0xAC])})
        tx = {"version": 0, "inputs": ins, "outputs": outs, "lock_time": 0, "subnetwork_id": SUBNET_NATIVE, "gas": 0, "payload": b"", "mass": 0}
        tx["mass"] = storage_mass([(e["amount"], len(e["script"])) for e in ents], [(o["value"], 34) for o in outs], storage_mass_parameter)
        for idx, (kind, ks, redeem) in enumerate(meta):
            ecdsa = kind in (KIND_P2PK_ECDSA, KIND_MS_ECDSA)
            msg = sighash_all(tx, ents, idx, ecdsa)
            if kind in (KIND_P2PK, KIND_P2PK_ECDSA):
                tx["inputs"][idx]["sigscript"] = bytes([0x41]) + dag._sign(ks[0], msg, ecdsa) + bytes([SIGHASH_ALL])
            else:
                pair = sorted(int(x) for x in rng.choice(3, size=2, replace=False))
                tx["inputs"][idx]["sigscript"] = b"".join(bytes([0x41]) + dag._sign(ks[p], msg, ecdsa) + bytes([SIGHASH_ALL]) for p in pair) + _push(redeem)
        txs.append(tx)
    return np.frombuffer(b"".join(fund_keys), dtype=np.uint8).reshape(-1, 36).copy(), fund_entries, txs


def entries_to_arrays(entries):
    """list of entry dicts -> (ENTRY_DTYPE array, byte arena) for kgv_utxo_apply_diff"""
    from .txbatch import ENTRY_DTYPE
    arr = np.zeros(len(entries), dtype=ENTRY_DTYPE)
    arena = bytearray()
    for i, e in enumerate(entries):
        arr[i]["amount"], arr[i]["block_daa_score"], arr[i]["spk_version"], arr[i]["is_coinbase"] = e["amount"], e.get("block_daa_score", 0), e["spk_version"], 1 if e.get("is_coinbase") else 0
        arr[i]["script_off"], arr[i]["script_len"] = len(arena), len(e["script"])
        arena.extend(e["script"])
    return arr, np.frombuffer(bytes(arena) + bytes(8), dtype=np.uint8).copy()


# ----------------------------------------------------------------------------------------------------
# fast generator (tools/simgen/kgv_simgen.cpp): the same shapes, ~100x faster, emitted directly in the flat batch layout
# ----------------------------------------------------------------------------------------------------
class FastDag:
    """Seeded C++ generator of a linearised chain of blocks of signed transactions (configs 1, 3, 4, 5 of BASELINE.json at
    their stated sizes).  generate(n_blocks, tpb) appends blocks; take() returns everything emitted since the last take()
    as (TxBatch, block_first_tx (n+1,), pov (n,)) and starts a new batch (spendable outputs carry over)."""

    def __init__(self, seed=0x6B61737061, n_keys=1024, n_nonces=4096, storage_mass_parameter=DEFAULT_STORAGE_MASS_PARAMETER,
                 coinbase_maturity=DEFAULT_COINBASE_MATURITY, mix=(1.0, 0.0, 0.0, 0.0), frac_two_inputs=0.5, frac_invalid=0.0,
                 coinbase_outputs=8, subsidy=50_000_000_000):
        import ctypes
        import os
        import subprocess
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        src = os.path.join(here, "tools", "simgen", "kgv_simgen.cpp")
        lib = os.path.join(here, "tools", "simgen", "libkgv_simgen.so")
        if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib, src], check=True, capture_output=True)
        self._ct = ctypes
        self._lib = ctypes.CDLL(lib)
        self._lib.sg_create.restype = ctypes.c_void_p
        keys = W.ScalarPointPool(n_keys, seed, b"sim-keys")
        nonces = W.ScalarPointPool(n_nonces, seed, b"sim-nonces")
        self.keys, self.nonces = keys, nonces
        self.C, self.maturity = storage_mass_parameter, coinbase_maturity

        class Cfg(ctypes.Structure):
            _fields_ = [("seed", ctypes.c_uint64), ("n_keys", ctypes.c_uint32), ("n_nonces", ctypes.c_uint32), ("smp", ctypes.c_uint64), ("maturity", ctypes.c_uint64),
                        ("mix", ctypes.c_double * 4), ("f2", ctypes.c_double), ("finv", ctypes.c_double), ("cbo", ctypes.c_uint32), ("pad_", ctypes.c_uint32),
                        ("subsidy", ctypes.c_uint64)]
        cfg = Cfg(seed, n_keys, n_nonces, storage_mass_parameter, coinbase_maturity, (ctypes.c_double * 4)(*[float(m) for m in mix]), frac_two_inputs, frac_invalid,
                  coinbase_outputs, 0, subsidy)
        be = lambda ints: b"".join(int(x).to_bytes(32, "big") for x in ints)
        self._h = ctypes.c_void_p(self._lib.sg_create(ctypes.byref(cfg), be(keys.scalars), b"".join(keys.xs), be(nonces.scalars),
                                                      be(pow(k, -1, N) for k in nonces.scalars), b"".join(nonces.xs)))

    def close(self):
        if self._h:
            self._lib.sg_destroy(self._h)
            self._h = None

    def generate(self, n_blocks, txs_per_block):
        self._lib.sg_generate(self._h, int(n_blocks), int(txs_per_block))

    def counts(self):
        c = (self._ct.c_uint64 * 8)()
        self._lib.sg_counts(self._h, c)
        return dict(zip(("n_txs", "n_inputs", "n_outputs", "n_bytes", "n_blocks", "n_signatures", "n_invalid", "n_utxos"), (int(x) for x in c)))

    def take(self):
        from .txbatch import INPUT_DTYPE, OUTPUT_DTYPE, TX_DTYPE, TxBatch
        c = self.counts()
        txs = np.zeros(c["n_txs"], dtype=TX_DTYPE)
        ins = np.zeros(c["n_inputs"], dtype=INPUT_DTYPE)
        outs = np.zeros(c["n_outputs"], dtype=OUTPUT_DTYPE)
        arena = np.zeros(c["n_bytes"] + 16, dtype=np.uint8)
        first = np.zeros(c["n_blocks"] + 1, dtype=np.uint32)
        pov = np.zeros(c["n_blocks"], dtype=np.uint64)
        p = lambda a: a.ctypes.data_as(self._ct.c_void_p)
        self._lib.sg_copy(self._h, p(txs), p(ins), p(outs), p(arena), p(first), p(pov))
        self._lib.sg_reset_output(self._h)
        return TxBatch(txs, ins, outs, None, arena), first, pov
