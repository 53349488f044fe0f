"""GPU-less unit tests of the DEVICE code: tests/hostsim compiles the CUDA headers with g++ using the
portable bodies of the PTX primitives, so the field/scalar/group logic, the GLV split, the window
recoding and both verification cores are exercised on the CPU against pyref and the C oracle.
(The PTX bodies themselves are covered by the -m gpu parity tests.)"""
import ctypes
import os
import random
import subprocess

import pytest

import pyref

HS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
P, N = pyref.P, pyref.N
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def _build(name):
    src, out = os.path.join(HS, name + ".cpp"), os.path.join(HS, "lib" + name + ".so")
    hdrs = [os.path.join(HS, "..", "..", "rusty_kaspa_b200", "csrc", f) for f in ("kgv_arith.cuh", "kgv_secp.cuh", "kgv_sha256.cuh", "kgv_verify.cuh", "kgv_u3072.cuh", "kgv_blake2b.cuh", "kgv_muhash.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(h) > os.path.getmtime(out) for h in hdrs + [src]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return ctypes.CDLL(out)


@pytest.fixture(scope="module")
def arith():
    return _build("hostsim_arith")


@pytest.fixture(scope="module")
def secp():
    return _build("hostsim_secp")


def le(x, n=32):
    return x.to_bytes(n, "little")


def edge(rnd):
    c = rnd.random()
    if c < 0.3:
        return rnd.randrange(2**256)
    if c < 0.4:
        return P + rnd.randrange(0, 2**32 + 977)  # non-canonical band
    if c < 0.5:
        return rnd.randrange(0, 2**33)
    if c < 0.6:
        return P - rnd.randrange(1, 2**33)
    if c < 0.7:
        return 2**256 - 1 - rnd.randrange(0, 5)
    if c < 0.8:
        return (2**256 - 1) ^ (1 << rnd.randrange(256))
    if c < 0.85:
        return rnd.randrange(2**96) | ((2**160 - 1) << 96)  # upper limbs all ones: the rare carry-propagation path
    if c < 0.9:
        return rnd.choice([0, 1, P, P - 1, P + 1, 2**255, 2**224 - 1, 0xFFFFFFFF << (32 * rnd.randrange(8))])
    return (rnd.randrange(2**256) | (0xFFFFFFFFFFFFFFFF << (64 * rnd.randrange(3)))) & (2**256 - 1)


def test_field_arithmetic(arith):
    rnd = random.Random(7)
    o = ctypes.create_string_buffer(64)
    iv = lambda: int.from_bytes(o.raw[:32], "little")
    for _ in range(20000):
        a, b = edge(rnd), edge(rnd)
        arith.hs_mul_wide(le(a), le(b), o)
        assert int.from_bytes(o.raw, "little") == a * b
        arith.hs_sqr_wide(le(a), o)
        assert int.from_bytes(o.raw, "little") == a * a
        arith.hs_fe_mul(le(a), le(b), o)
        assert iv() % P == a * b % P
        arith.hs_fe_sqr(le(a), o)
        assert iv() % P == a * a % P
        arith.hs_fe_add(le(a), le(b), o)
        assert iv() % P == (a + b) % P
        arith.hs_fe_sub(le(a), le(b), o)
        assert iv() % P == (a - b) % P
        arith.hs_fe_mul8(le(a), o)
        assert iv() % P == 8 * a % P
        arith.hs_fe_mul3(le(a), o)
        assert iv() % P == 3 * a % P
        arith.hs_fe_norm(le(a), o)
        assert iv() == a % P
        assert bool(arith.hs_fe_is_zero(le(a))) == (a % P == 0)
    for _ in range(100):
        a = edge(rnd)
        arith.hs_fe_inv(le(a), o)
        assert iv() % P == pow(a, P - 2, P)
        ok = arith.hs_fe_sqrt(le(a), o)
        assert bool(ok) == (pow(a % P, (P - 1) // 2, P) in (0, 1))
        if ok:
            assert iv() ** 2 % P == a % P


def test_reduce_wide_rare_carry_paths(arith):
    """512-bit inputs crafted so that the second fold's carry runs through limb 3, through limbs 4..7, and past
    2^256 (the normally untaken branches of fe_reduce_wide)."""
    rnd = random.Random(11)
    C = 2**32 + 977
    o = ctypes.create_string_buffer(32)
    hit = [0, 0, 0]
    for it in range(3000):
        hi = rnd.randrange(2**255, 2**256) if it % 2 else rnd.randrange(1, 2**256)
        kind = it % 3
        if kind == 0:    # carry leaves limb 3 only
            xs = (rnd.randrange(2**128) << 128) | (2**128 - 1 - rnd.randrange(2**20))
        elif kind == 1:  # ripples through some of limbs 4..7
            k = rnd.randrange(1, 4)
            xs = (rnd.randrange(2**(32 * (4 - k))) << (32 * (4 + k))) | (2**(32 * (4 + k)) - 1 - rnd.randrange(2**20))
        else:            # wraps past 2^256
            xs = 2**256 - 1 - rnd.randrange(2**20)
        lo = (xs - hi * C) % 2**256
        t = lo + (hi << 256)
        arith.hs_fe_reduce_wide(t.to_bytes(64, "little"), o)
        got = int.from_bytes(o.raw, "little")
        assert got % P == t % P, (hex(t), hex(got))
        top = (lo + hi * C) >> 256
        s2 = xs + top * C
        hit[0] += (xs % 2**128 + top * C) >> 128 > 0
        hit[1] += kind == 1 and (xs % 2**160 + top * C) >> 160 > 0
        hit[2] += s2 >> 256 > 0
    assert min(hit) > 100, hit


def limbs(x, n=8):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(a):
    return sum(int(v) << (32 * i) for i, v in enumerate(a))


def test_glv_split_and_scalars(secp):
    rnd = random.Random(11)
    specials = [0, 1, 2, N - 1, N - 2, LAMBDA, N - LAMBDA, (N - 1) // 2, 3, 4, 5]
    for it in range(5000):
        k = specials[it] if it < len(specials) else rnd.randrange(N)
        k1, k2 = (ctypes.c_uint32 * 5)(), (ctypes.c_uint32 * 5)()
        n1, n2 = ctypes.c_int(), ctypes.c_int()
        secp.hs_glv_split(limbs(k), k1, ctypes.byref(n1), k2, ctypes.byref(n2))
        a = val(k1) * (-1 if n1.value else 1)
        b = val(k2) * (-1 if n2.value else 1)
        assert (a + b * LAMBDA - k) % N == 0 and abs(a) < 2**128 and abs(b) < 2**128
    r = (ctypes.c_uint32 * 8)()
    for it in range(200):
        a, b = rnd.randrange(1, N), rnd.randrange(N)
        secp.hs_sc_mul(limbs(a), limbs(b), r)
        assert val(r) == a * b % N
        if it < 10:
            secp.hs_sc_inv(limbs(a), r)
            assert val(r) == pow(a, -1, N)


def test_generator_table_entries_and_hashes(secp):
    xy = (ctypes.c_uint32 * 16)()
    for v, w in [(1, 0), (2, 0), (65535, 0), (1, 1), (40000, 1), (12345, 0)]:
        secp.hs_gtab_entry(v, w, xy)
        pt = pyref.pt_mul(v * (2**128 if w else 1), pyref.G)
        assert val(xy[:8]) == pt[0] and val(xy[8:]) == pt[1]
    rnd = random.Random(3)
    rb = lambda n: bytes(rnd.getrandbits(8) for _ in range(n))
    o = ctypes.create_string_buffer(32)
    r_, p_, m_ = rb(32), rb(32), rb(32)
    secp.hs_sha_challenge(r_, p_, m_, o)
    assert o.raw == pyref.tagged_hash("BIP0340/challenge", r_ + p_ + m_)
    secp.hs_ecdsa_wrap(m_, o)
    assert o.raw == pyref.sha256_domain(b"TransactionSigningHashECDSA", m_)


def test_verify_cores_match_oracle(secp, oracle):
    import numpy as np
    from rusty_kaspa_b200 import workload as W
    pk, msg, sig, kind = W.schnorr_triples(120, seed=21, n_keys=8, n_nonces=8, frac_bitflip=0.25, frac_adversarial=0.35)
    seen = set()
    for i in range(len(pk)):
        got = secp.hs_schnorr_verify(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes())
        assert got == oracle.ok_schnorr_verify(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes()), (i, kind[i])
        seen.add(got)
    assert seen == {0, 1, 2}
    pk, msg, sig, kind = W.ecdsa_triples(60, seed=22, n_keys=8, n_nonces=8, frac_bitflip=0.25, frac_adversarial=0.45)
    seen = set()
    for i in range(len(pk)):
        got = secp.hs_ecdsa_verify(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes())
        assert got == oracle.ok_ecdsa_verify(pk[i].tobytes(), msg[i].tobytes(), sig[i].tobytes()), (i, kind[i])
        seen.add(got)
    assert seen == {0, 1, 2, 3}


def test_u3072_field_and_element_expansion():
    """MuHash field (modulo 2^3072 - 1103717) and ChaCha20 element expansion of kgv_u3072.cuh against Python integers:
    products incl. values >= p, the crafted second-level wrap of the fold, the block-transposed layout, canonicalisation."""
    L = _build("hostsim_u3072")
    MP = pyref.MUHASH_P
    arr = lambda v, n=96: (ctypes.c_uint32 * n)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
    rnd = random.Random(3)

    def big():
        c = rnd.random()
        if c < 0.5:
            return rnd.getrandbits(3072)
        if c < 0.7:
            return rnd.choice([0, 1, 2, MP - 1, MP, MP + 1, 2**3072 - 1, 2**3072 - 2, 1103717, 1103716, 2**3071])
        if c < 0.85:
            return (2**3072 - 1) ^ rnd.getrandbits(40)
        return (2**3072 - 1) ^ (rnd.getrandbits(64) << (32 * rnd.randrange(94)))

    for _ in range(300):
        a, b = big(), big()
        r, w = (ctypes.c_uint32 * 96)(), (ctypes.c_uint32 * 192)()
        L.hs_u3072_mul_mod(arr(a), arr(b), r, w)
        assert val(w) == a * b
        assert val(r) % MP == a * b % MP and val(r) < 2**3072
        c = (ctypes.c_uint32 * 96)()
        L.hs_u3072_canonical(r, c)
        assert val(c) == a * b % MP
        L.hs_u3072_canonical(arr(a), c)
        assert val(c) == a % MP
    for it in range(200):  # low half of (lo + hi * PRIME_DIFF) all ones: the extra fold ripples to the top
        hi = rnd.getrandbits(3072) if it % 2 else 2**3072 - 1 - rnd.getrandbits(30)
        lo = ((2**3072 - 1) - rnd.getrandbits(20) - hi * 1103717) % 2**3072
        r = (ctypes.c_uint32 * 96)()
        L.hs_u3072_fold(arr(lo + (hi << 3072), 192), r)
        assert val(r) % MP == (lo + (hi << 3072)) % MP and val(r) < 2**3072
    S, SS = 5, 3  # strided (block-transposed) arrays
    vals = [rnd.getrandbits(3072) for _ in range(S)]
    A = (ctypes.c_uint32 * (96 * S))()
    for e, v in enumerate(vals):
        for blk in range(12):
            for j in range(8):
                A[(blk * S + e) * 8 + j] = (v >> (32 * (8 * blk + j))) & 0xFFFFFFFF
    scratch = (ctypes.c_uint32 * (192 * SS))()
    sz = ctypes.c_size_t
    L.hs_u3072_mul_mod_strided(A, sz(S), sz(1), sz(3), sz(4), scratch, sz(SS), sz(2))
    get = lambda e: sum(A[(blk * S + e) * 8 + j] << (32 * (8 * blk + j)) for blk in range(12) for j in range(8))
    assert get(4) % MP == vals[1] * vals[3] % MP and [get(e) for e in range(4)] == vals[:4]
    for n in (0, 1, 13, 127, 128, 129, 300):  # the two keyed domains (key block pending as the last / a middle block)
        d = bytes(rnd.randrange(256) for _ in range(n))
        o32 = ctypes.create_string_buffer(32)
        L.hs_muhash_domain_hash(0, d, ctypes.c_size_t(n), o32)
        assert o32.raw == pyref.blake2b_keyed(b"MuHashElement", d)
        L.hs_muhash_domain_hash(1, d, ctypes.c_size_t(n), o32)
        assert o32.raw == pyref.blake2b_keyed(b"MuHashFinalize", d)
    for _ in range(50):
        d = bytes(rnd.randrange(256) for _ in range(rnd.randrange(0, 100)))
        o = (ctypes.c_uint32 * 96)()
        L.hs_muhash_expand(pyref.blake2b_keyed(b"MuHashElement", d), o)
        assert val(o) == pyref.muhash_element(d)
