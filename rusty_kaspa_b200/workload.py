"""Synthetic workloads for the validation hot path (BASELINE.json configs).

Signature triples are produced WITHOUT any elliptic-curve library and without touching oracle/:
a seeded pool of keys (d_i, P_i = d_i*G) and a pool of nonces (k_j, R_j = k_j*G) are built with
incremental affine additions on Python integers, then every BIP-340 signature is just
    e = H_tag(R_j.x || P_i.x || m) mod n ,  s = k_j + e*d_i mod n
(scalar arithmetic only).  Nonce reuse is irrelevant here: these are benchmark inputs, the
verifier neither caches nor batches across signatures, and every triple verifies independently.
This mirrors what simpa's miner does with `sign()` (simpa/src/simulator/miner.py:138-172 shape:
valid Schnorr P2PK spends), restated as a seeded generator (simpa itself uses thread_rng).
"""
import hashlib
import struct

import numpy as np

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8

KIND_VALID, KIND_BITFLIP, KIND_ADVERSARIAL = 0, 1, 2


def _add(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (b[1] - a[1]) * pow(b[0] - a[0], -1, P) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def _mul(k, pt):
    r = None
    while k:
        if k & 1:
            r = _add(r, pt)
        pt = _add(pt, pt)
        k >>= 1
    return r


def _seed_int(seed, label):
    return int.from_bytes(hashlib.sha256(b"kgv-workload|%s|%d" % (label, seed)).digest(), "big") % (N - 1) + 1


class ScalarPointPool:
    """K pairs (s_i, S_i = s_i*G) with even-y points: s_i = +-(s_0 + i*delta)."""

    def __init__(self, count, seed, label):
        s0 = _seed_int(seed, label + b"|base")
        delta = _seed_int(seed, label + b"|step")
        step_pt = _mul(delta, (GX, GY))
        pt = _mul(s0, (GX, GY))
        self.scalars = []
        self.xs = []  # 32-byte big-endian x coordinates
        s = s0
        for _ in range(count):
            if pt[1] & 1:
                self.scalars.append(N - s)
            else:
                self.scalars.append(s)
            self.xs.append(pt[0].to_bytes(32, "big"))
            pt = _add(pt, step_pt)
            s = (s + delta) % N
        self.count = count


_TAG_CHALLENGE = hashlib.sha256(hashlib.sha256(b"BIP0340/challenge").digest() * 2)


def _challenge(rx, px, m):
    h = _TAG_CHALLENGE.copy()
    h.update(rx)
    h.update(px)
    h.update(m)
    return int.from_bytes(h.digest(), "big") % N


def _non_residue_x(rng):
    """An x < p with x^3+7 a quadratic non-residue (no curve point)."""
    while True:
        x = int.from_bytes(rng.bytes(32), "big") % P
        if pow((pow(x, 3, P) + 7) % P, (P - 1) // 2, P) == P - 1:
            return x


def schnorr_triples(n, seed=0x6B61737061, n_keys=65536, n_nonces=65536, frac_bitflip=0.01, frac_adversarial=0.01, pools=None):
    """BASELINE config 2: n (pk, msg, sig) triples, SoA uint8 arrays + a per-item `kind` array.

    ~98 % valid, ~1 % single-bit corruptions (uniform over sig/msg/pk), ~1 % adversarial encodings
    (r >= p, s >= n, x(pk) >= p, off-curve pk, s = 0, R = infinity, odd-y R, all-zero).
    Items of kind 0 are valid by construction; everything else is NOT valid (verdict 0 or 2).
    """
    rng = np.random.default_rng(seed)
    n_keys = min(n_keys, max(1, n))
    n_nonces = min(n_nonces, max(1, n))
    if pools is None:
        pools = (ScalarPointPool(n_keys, seed, b"keys"), ScalarPointPool(n_nonces, seed, b"nonces"))
    keys, nonces = pools
    msgs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ki = rng.integers(0, keys.count, size=n)
    ni = rng.integers(0, nonces.count, size=n)
    u = rng.random(n)
    kind = np.zeros(n, dtype=np.uint8)
    kind[u < frac_bitflip + frac_adversarial] = KIND_ADVERSARIAL
    kind[u < frac_bitflip] = KIND_BITFLIP
    pk = np.empty((n, 32), dtype=np.uint8)
    sig = np.empty((n, 64), dtype=np.uint8)
    msg_bytes = msgs.tobytes()
    pk_out = bytearray(32 * n)
    sig_out = bytearray(64 * n)
    for i in range(n):
        d = keys.scalars[ki[i]]
        px = keys.xs[ki[i]]
        k = nonces.scalars[ni[i]]
        rx = nonces.xs[ni[i]]
        m = msg_bytes[32 * i:32 * i + 32]
        e = _challenge(rx, px, m)
        s = (k + e * d) % N
        if kind[i] == KIND_ADVERSARIAL:
            c = int(rng.integers(0, 8))
            if c == 0:    # r >= p
                rx = (P + int(rng.integers(0, 2**32 + 977))).to_bytes(32, "big")
            elif c == 1:  # s >= n
                s = N + int(rng.integers(0, 2**62))
            elif c == 2:  # x(pk) >= p
                px = (P + int(rng.integers(0, 2**32 + 977))).to_bytes(32, "big")
            elif c == 3:  # off-curve pk
                px = _non_residue_x(rng).to_bytes(32, "big")
            elif c == 4:  # s = 0
                s = 0
            elif c == 5:  # R = s*G - e*P = infinity
                s = e * d % N
            elif c == 6:  # R has the right x but odd y
                s = (N - k + e * d) % N
            else:         # all-zero signature
                rx = bytes(32)
                s = 0
        pk_out[32 * i:32 * i + 32] = px
        sig_out[64 * i:64 * i + 32] = rx
        sig_out[64 * i + 32:64 * i + 64] = s.to_bytes(32, "big")
    pk[:] = np.frombuffer(bytes(pk_out), dtype=np.uint8).reshape(n, 32)
    sig[:] = np.frombuffer(bytes(sig_out), dtype=np.uint8).reshape(n, 64)
    # single-bit corruptions, uniform over the 128 bytes of the triple
    flip_idx = np.nonzero(kind == KIND_BITFLIP)[0]
    for i in flip_idx:
        pos = int(rng.integers(0, 128 * 8))
        byte, bit = pos >> 3, pos & 7
        if byte < 32:
            pk[i, byte] ^= 1 << bit
        elif byte < 64:
            msgs[i, byte - 32] ^= 1 << bit
        else:
            sig[i, byte - 64] ^= 1 << bit
    return pk, msgs, sig, kind


def tile_triples(pk, msg, sig, kind, n):
    """Repeat a base set of triples up to n items (keeps generation time bounded for huge batches)."""
    reps = (n + len(pk) - 1) // len(pk)
    f = lambda a: np.ascontiguousarray(np.tile(a, (reps, 1))[:n])
    return f(pk), f(msg), f(sig), np.tile(kind, reps)[:n].copy()


def ecdsa_triples(n, seed=0x6B61737061, n_keys=65536, n_nonces=65536, frac_bitflip=0.01, frac_adversarial=0.01, pools=None):
    """ECDSA (pk33, msg32, sig64) triples for BASELINE config 4, same pool technique:
    r = x(k_j*G) mod n, s = k_j^-1 (m + r*d_i) mod n, normalised to low S.
    Adversarial kinds: high S, r >= n, s >= n, r = 0, s = 0, bad key tag, x(pk) >= p, off-curve pk."""
    rng = np.random.default_rng(seed ^ 0xEC)
    n_keys = min(n_keys, max(1, n))
    n_nonces = min(n_nonces, max(1, n))
    if pools is None:
        pools = (ScalarPointPool(n_keys, seed, b"keys"), ScalarPointPool(n_nonces, seed, b"nonces"))
    keys, nonces = pools
    kinv = getattr(nonces, "inv", None)
    if kinv is None:
        kinv = nonces.inv = [pow(k, -1, N) for k in nonces.scalars]
    msgs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ki = rng.integers(0, keys.count, size=n)
    ni = rng.integers(0, nonces.count, size=n)
    u = rng.random(n)
    kind = np.zeros(n, dtype=np.uint8)
    kind[u < frac_bitflip + frac_adversarial] = KIND_ADVERSARIAL
    kind[u < frac_bitflip] = KIND_BITFLIP
    msg_bytes = msgs.tobytes()
    pk_out = bytearray(33 * n)
    sig_out = bytearray(64 * n)
    half = N // 2
    for i in range(n):
        d = keys.scalars[ki[i]]
        pkb = b"\x02" + keys.xs[ki[i]]  # pool points have even y
        r = int.from_bytes(nonces.xs[ni[i]], "big") % N
        m = int.from_bytes(msg_bytes[32 * i:32 * i + 32], "big") % N
        s = kinv[ni[i]] * (m + r * d) % N
        if s > half:
            s = N - s
        if kind[i] == KIND_ADVERSARIAL:
            c = int(rng.integers(0, 8))
            if c == 0:
                s = N - s                       # high S: well-formed, rejected by verify
            elif c == 1:
                r = N + int(rng.integers(0, 2**62))  # r >= n: parse error
            elif c == 2:
                s = N + int(rng.integers(0, 2**62))  # s >= n: parse error
            elif c == 3:
                r = 0
            elif c == 4:
                s = 0
            elif c == 5:
                pkb = bytes([int(rng.choice([0, 1, 4, 6, 7, 0xFF]))]) + pkb[1:]
            elif c == 6:
                pkb = pkb[:1] + (P + int(rng.integers(0, 2**32 + 977))).to_bytes(32, "big")
            else:
                pkb = pkb[:1] + _non_residue_x(rng).to_bytes(32, "big")
        pk_out[33 * i:33 * i + 33] = pkb
        sig_out[64 * i:64 * i + 32] = r.to_bytes(32, "big")
        sig_out[64 * i + 32:64 * i + 64] = s.to_bytes(32, "big")
    pk = np.frombuffer(bytes(pk_out), dtype=np.uint8).reshape(n, 33).copy()
    sig = np.frombuffer(bytes(sig_out), dtype=np.uint8).reshape(n, 64).copy()
    for i in np.nonzero(kind == KIND_BITFLIP)[0]:
        pos = int(rng.integers(0, 129 * 8))
        byte, bit = pos >> 3, pos & 7
        if byte < 33:
            pk[i, byte] ^= 1 << bit
        elif byte < 65:
            msgs[i, byte - 33] ^= 1 << bit
        else:
            sig[i, byte - 65] ^= 1 << bit
    return pk, msgs, sig, kind


# ------------------------------------------------------------------------------------------------
# random transactions for the hashing kernels (shape-only: arbitrary scripts / payloads)
# ------------------------------------------------------------------------------------------------
def random_transactions(n, seed=1, max_inputs=4, max_outputs=4, max_script=80, max_payload=300):
    """n arbitrary transactions + a populated entry per input (ragged sizes, empty scripts/payloads,
    coinbase / native / other subnetworks, non-zero mass) for tx-id / tx-hash / sighash parity tests."""
    rng = np.random.default_rng(seed)
    rb = lambda k: rng.integers(0, 256, size=int(k), dtype=np.uint8).tobytes()
    txs, entries = [], []
    for t in range(n):
        kind = int(rng.integers(0, 10))
        subnet = bytes(20) if kind < 7 else (bytes([1]) + bytes(19) if kind == 7 else rb(20))
        n_in = 0 if (kind == 7 and rng.random() < 0.5) else int(rng.integers(0 if kind >= 7 else 1, max_inputs + 1))
        n_out = int(rng.integers(0, max_outputs + 1))
        ins = [{"txid": rb(32), "index": int(rng.integers(0, 2**32)), "sigscript": rb(rng.integers(0, max_script)),
                "sequence": int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2)), "sig_op_count": int(rng.integers(0, 256))} for _ in range(n_in)]
        outs = [{"value": int(rng.integers(0, 2**62)), "spk_version": int(rng.integers(0, 3)), "script": rb(rng.integers(0, max_script))} for _ in range(n_out)]
        payload = b"" if rng.random() < 0.5 else rb(rng.integers(1, max_payload))
        txs.append({"version": int(rng.integers(0, 3)), "inputs": ins, "outputs": outs, "lock_time": int(rng.integers(0, 2**62)),
                    "subnetwork_id": subnet, "gas": int(rng.integers(0, 2**40)), "payload": payload,
                    "mass": 0 if rng.random() < 0.5 else int(rng.integers(1, 2**40))})
        entries.append([{"amount": int(rng.integers(0, 2**62)), "spk_version": int(rng.integers(0, 2)), "script": rb(rng.integers(0, max_script)),
                         "block_daa_score": int(rng.integers(0, 2**40)), "is_coinbase": bool(rng.integers(0, 2))} for _ in range(n_in)])
    return txs, entries
