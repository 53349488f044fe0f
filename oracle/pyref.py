"""ORACLE — TEST INFRASTRUCTURE ONLY (second, independent restatement).

Pure-Python big-int twin of oracle/*.c.  It exists so that the two restatements can
be cross-checked against each other on random + adversarial inputs for the cases the
reference's own vectors do not pin (ECDSA verdicts, Schnorr edge encodings).
Only tests/ may import it.  Slow (tens of ms per verify): small cases only.

Follows:
  crypto/txscript/src/lib.rs:574-643            check_schnorr_signature / check_ecdsa_signature
  consensus/core/src/hashing/tx.rs:16-107        tx id / tx hash encoding
  consensus/core/src/hashing/sighash.rs:140-277  sighash
  crypto/hashes/src/hashers.rs:23-106            keyed BLAKE2b-256 / domain SHA-256 hashers
and, for the curve arithmetic the reference delegates to libsecp256k1 (not vendored),
the published BIP-340 and SEC 1 algorithms.
"""
import hashlib
import struct

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
G = (GX, GY)

INVALID, VALID, PK_PARSE_ERR, SIG_PARSE_ERR = 0, 1, 2, 3


# ---------------------------------------------------------------- curve
def pt_add(a, b):
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


def pt_mul(k, pt):
    r = None
    k %= N
    while k:
        if k & 1:
            r = pt_add(r, pt)
        pt = pt_add(pt, pt)
        k >>= 1
    return r


def lift_x(x, odd=False):
    """Returns the curve point with this x and the requested y parity, or None."""
    if x >= P:
        return None
    y2 = (pow(x, 3, P) + 7) % P
    y = pow(y2, (P + 1) // 4, P)
    if y * y % P != y2:
        return None
    if (y & 1) != int(odd):
        y = P - y
    return (x, y)


def tagged_hash(tag, data):
    th = hashlib.sha256(tag.encode()).digest()
    return hashlib.sha256(th + th + data).digest()


# ---------------------------------------------------------------- BIP-340 / ECDSA
def schnorr_verify(pk32, msg32, sig64):
    pt = lift_x(int.from_bytes(pk32, "big"))
    if pt is None:
        return PK_PARSE_ERR
    r = int.from_bytes(sig64[:32], "big")
    s = int.from_bytes(sig64[32:], "big")
    if r >= P or s >= N:
        return INVALID
    e = int.from_bytes(tagged_hash("BIP0340/challenge", sig64[:32] + pk32 + msg32), "big") % N
    R = pt_add(pt_mul(s, G), pt_mul(N - e, pt))
    if R is None or (R[1] & 1) or R[0] != r:
        return INVALID
    return VALID


def schnorr_pubkey(sk32):
    d = int.from_bytes(sk32, "big")
    assert 0 < d < N
    return pt_mul(d, G)[0].to_bytes(32, "big")


def schnorr_sign(sk32, msg32, aux=b"\x00" * 32):
    d0 = int.from_bytes(sk32, "big")
    assert 0 < d0 < N
    Pt = pt_mul(d0, G)
    d = d0 if Pt[1] % 2 == 0 else N - d0
    px = Pt[0].to_bytes(32, "big")
    t = bytes(a ^ b for a, b in zip(d.to_bytes(32, "big"), tagged_hash("BIP0340/aux", aux)))
    k0 = int.from_bytes(tagged_hash("BIP0340/nonce", t + px + msg32), "big") % N
    assert k0
    R = pt_mul(k0, G)
    k = k0 if R[1] % 2 == 0 else N - k0
    rx = R[0].to_bytes(32, "big")
    e = int.from_bytes(tagged_hash("BIP0340/challenge", rx + px + msg32), "big") % N
    return rx + ((k + e * d) % N).to_bytes(32, "big")


def ecdsa_verify(pk33, msg32, sig64):
    if pk33[0] not in (2, 3):
        return PK_PARSE_ERR
    Q = lift_x(int.from_bytes(pk33[1:], "big"), odd=(pk33[0] == 3))
    if Q is None:
        return PK_PARSE_ERR
    r = int.from_bytes(sig64[:32], "big")
    s = int.from_bytes(sig64[32:], "big")
    if r >= N or s >= N:
        return SIG_PARSE_ERR
    if s > N // 2 or r == 0 or s == 0:
        return INVALID
    m = int.from_bytes(msg32, "big") % N
    si = pow(s, -1, N)
    R = pt_add(pt_mul(m * si % N, G), pt_mul(r * si % N, Q))
    if R is None or R[0] % N != r:
        return INVALID
    return VALID


def ecdsa_pubkey(sk32):
    d = int.from_bytes(sk32, "big")
    Pt = pt_mul(d, G)
    return bytes([2 + (Pt[1] & 1)]) + Pt[0].to_bytes(32, "big")


# ---------------------------------------------------------------- hashing
def blake2b_keyed(domain, data):
    return hashlib.blake2b(data, digest_size=32, key=domain).digest()


def sha256_domain(domain, data):
    return hashlib.sha256(hashlib.sha256(domain).digest() + data).digest()


SUBNETWORK_NATIVE = bytes(20)
SUBNETWORK_COINBASE = bytes([1]) + bytes(19)


def _varbytes(b):
    return struct.pack("<Q", len(b)) + b


def _enc_output(o):
    return struct.pack("<QH", o["value"], o["spk_version"]) + _varbytes(o["script"])


def _enc_tx(tx, exclude_sigscript, exclude_mass):
    """consensus/core/src/hashing/tx.rs:45-107"""
    b = struct.pack("<HQ", tx["version"], len(tx["inputs"]))
    for i in tx["inputs"]:
        b += i["txid"] + struct.pack("<I", i["index"])
        if exclude_sigscript:
            b += _varbytes(b"")
        else:
            b += _varbytes(i["sigscript"]) + bytes([i["sig_op_count"]])
        b += struct.pack("<Q", i["sequence"])
    b += struct.pack("<Q", len(tx["outputs"]))
    for o in tx["outputs"]:
        b += _enc_output(o)
    b += struct.pack("<Q", tx["lock_time"]) + tx["subnetwork_id"] + struct.pack("<Q", tx["gas"]) + _varbytes(tx["payload"])
    if not exclude_mass and tx.get("mass", 0) > 0:
        b += struct.pack("<Q", tx["mass"])
    return b


def tx_id(tx):
    cb = tx["subnetwork_id"] == SUBNETWORK_COINBASE
    return blake2b_keyed(b"TransactionID", _enc_tx(tx, not cb, not cb))


def tx_hash(tx):
    return blake2b_keyed(b"TransactionHash", _enc_tx(tx, False, False))


SIGHASH_ALL, SIGHASH_NONE, SIGHASH_SINGLE, SIGHASH_ANYONECANPAY = 1, 2, 4, 0x80
ALLOWED_SIGHASH = (1, 2, 4, 0x81, 0x82, 0x84)


def sighash_schnorr(tx, entries, idx, hash_type):
    """consensus/core/src/hashing/sighash.rs:140-265. entries[i] = dict(amount, spk_version, script)"""
    H = lambda d: blake2b_keyed(b"TransactionSigningHash", d)
    Z = bytes(32)
    acp = bool(hash_type & SIGHASH_ANYONECANPAY)
    base = hash_type & 7
    prev = Z if acp else H(b"".join(i["txid"] + struct.pack("<I", i["index"]) for i in tx["inputs"]))
    seqs = Z if (acp or base in (SIGHASH_SINGLE, SIGHASH_NONE)) else H(b"".join(struct.pack("<Q", i["sequence"]) for i in tx["inputs"]))
    sops = Z if acp else H(bytes(i["sig_op_count"] for i in tx["inputs"]))
    if base == SIGHASH_NONE:
        outs = Z
    elif base == SIGHASH_SINGLE:
        outs = H(_enc_output(tx["outputs"][idx])) if idx < len(tx["outputs"]) else Z
    else:
        outs = H(b"".join(_enc_output(o) for o in tx["outputs"]))
    pay = Z if (tx["subnetwork_id"] == SUBNETWORK_NATIVE and not tx["payload"]) else H(_varbytes(tx["payload"]))
    i, e = tx["inputs"][idx], entries[idx]
    pre = (struct.pack("<H", tx["version"]) + prev + seqs + sops + i["txid"] + struct.pack("<I", i["index"])
           + struct.pack("<H", e["spk_version"]) + _varbytes(e["script"]) + struct.pack("<QQ", e["amount"], i["sequence"])
           + bytes([i["sig_op_count"]]) + outs + struct.pack("<Q", tx["lock_time"]) + tx["subnetwork_id"]
           + struct.pack("<Q", tx["gas"]) + pay + bytes([hash_type]))
    return H(pre)


def sighash_ecdsa(tx, entries, idx, hash_type):
    return sha256_domain(b"TransactionSigningHashECDSA", sighash_schnorr(tx, entries, idx, hash_type))


# ---------------------------------------------------------------------------------------------
# MuHash (crypto/muhash/src/lib.rs, u3072.rs; consensus/core/src/muhash.rs) — big-int twin of ok_muhash.c
# ---------------------------------------------------------------------------------------------
MUHASH_P = 2**3072 - 1103717


def _chacha20_block(key_words, counter):
    s = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [counter & 0xFFFFFFFF, counter >> 32, 0, 0]
    x = list(s)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 16) | (x[d] >> 16)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 12) | (x[b] >> 20)) & 0xFFFFFFFF
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] ^= x[a]; x[d] = ((x[d] << 8) | (x[d] >> 24)) & 0xFFFFFFFF
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] ^= x[c]; x[b] = ((x[b] << 7) | (x[b] >> 25)) & 0xFFFFFFFF

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & 0xFFFFFFFF for a, b in zip(x, s)])


def muhash_element(data):
    """lib.rs:161-166 data_to_element: keyed BLAKE2b -> ChaCha20Rng seed -> 384 keystream bytes, little-endian integer."""
    h = blake2b_keyed(b"MuHashElement", data)
    key = struct.unpack("<8I", h)
    return int.from_bytes(b"".join(_chacha20_block(key, c) for c in range(6)), "little")


class MuHash:
    def __init__(self):
        self.num, self.den = 1, 1

    def add_element(self, data):
        self.num = self.num * muhash_element(data) % MUHASH_P

    def remove_element(self, data):
        self.den = self.den * muhash_element(data) % MUHASH_P

    def combine(self, other):
        self.num = self.num * other.num % MUHASH_P
        self.den = self.den * other.den % MUHASH_P

    def serialize(self):
        self.num = self.num * pow(self.den, MUHASH_P - 2, MUHASH_P) % MUHASH_P
        self.den = 1
        return self.num.to_bytes(384, "little")

    def finalize(self):
        return blake2b_keyed(b"MuHashFinalize", self.serialize())


def utxo_element_bytes(txid, index, daa, amount, is_coinbase, spk_version, script):
    """consensus/core/src/muhash.rs:47-59 write_utxo"""
    return txid + struct.pack("<IQQ", index, daa, amount) + (b"\x01" if is_coinbase else b"\x00") + struct.pack("<HQ", spk_version, len(script)) + script


def merkle_root(hashes):
    """crypto/merkle/src/lib.rs:3-30, literally (array of 2*pot-1 optional nodes)"""
    if not hashes:
        return bytes(32)
    pot = 1
    while pot < len(hashes):
        pot *= 2
    nodes = list(hashes) + [None] * (2 * pot - 1 - len(hashes))
    off = pot
    for i in range(0, 2 * pot - 2, 2):
        nodes[off] = None if nodes[i] is None else blake2b_keyed(b"MerkleBranchHash", nodes[i] + (nodes[i + 1] if nodes[i + 1] is not None else bytes(32)))
        off += 1
    return nodes[-1]
