"""GPU parity: kgv_schnorr_verify (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import oracle_schnorr_batch
from rusty_kaspa_b200 import workload as W

pytestmark = pytest.mark.gpu


def test_generator_tables_match_reference_points(gpu_ctx):
    import pyref
    for which, v in [(0, 1), (0, 2), (0, 3), (0, 65535), (0, 12345), (1, 1), (1, 40000), (1, 65535)]:
        exp = pyref.pt_mul(v * (2**128 if which else 1), pyref.G)
        assert gpu_ctx.gtable_entry(which, v) == exp


def test_schnorr_parity_mixed_batch(gpu_ctx, oracle):
    pk, msg, sig, kind = W.schnorr_triples(20000, seed=1, n_keys=2048, n_nonces=2048, frac_bitflip=0.1, frac_adversarial=0.1)
    got = gpu_ctx.verify_schnorr_batch(pk, msg, sig)
    exp = oracle_schnorr_batch(oracle, pk, msg, sig)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first at {bad[:5]}: got {got[bad[:5]]} exp {exp[bad[:5]]} kind {kind[bad[:5]]}"
    assert (got[kind == 0] == 1).all()
    assert (got[kind != 0] != 1).all()


@pytest.mark.parametrize("n", [0, 1, 2, 31, 127, 128, 129, 1000])
def test_schnorr_ragged_sizes(gpu_ctx, oracle, n):
    pk, msg, sig, kind = W.schnorr_triples(max(n, 1), seed=n + 7, n_keys=16, n_nonces=16, frac_bitflip=0.2, frac_adversarial=0.2)
    pk, msg, sig = pk[:n], msg[:n], sig[:n]
    got = gpu_ctx.verify_schnorr_batch(pk, msg, sig, n=n)
    if n:
        assert (got == oracle_schnorr_batch(oracle, pk, msg, sig)).all()


def test_bitmap(gpu_ctx):
    rng = np.random.default_rng(3)
    for n in [1, 7, 8, 9, 1000, 4097]:
        st = rng.integers(0, 4, size=n, dtype=np.uint8)
        bm = gpu_ctx.status_to_bitmap(st)
        exp = np.packbits((st == 1).astype(np.uint8), bitorder="little")
        assert (bm == exp).all()


def test_bip340_test_vectors_on_the_gpu(gpu_ctx):
    """rows 0-14 of BIP-340's test-vectors.csv through kgv_schnorr_verify: verdicts incl. the parse-error class for unliftable keys;
    also embedded in a larger batch at every position of a warp so that the rare group-law paths (R at infinity) are taken by single lanes"""
    from golden_util import bip340_vectors
    pk, msg, sig, exp, comments = bip340_vectors()
    got = gpu_ctx.verify_schnorr_batch(pk, msg, sig)
    assert got.tolist() == exp, [(i, c) for i, (g, e, c) in enumerate(zip(got, exp, comments)) if g != e]
    fpk, fmsg, fsig, kind = W.schnorr_triples(512, seed=3, n_keys=32, n_nonces=32, frac_bitflip=0.0, frac_adversarial=0.0)
    for shift in range(0, 32, 5):
        P, M, S = fpk.copy(), fmsg.copy(), fsig.copy()
        pos = [(37 * i + shift) % 512 for i in range(15)]
        P[pos], M[pos], S[pos] = pk, msg, sig
        g = gpu_ctx.verify_schnorr_batch(P, M, S)
        assert g[pos].tolist() == exp and (np.delete(g, pos) == 1).all()
