"""GPU parity: kgv_ecdsa_verify (through the C ABI) vs the CPU oracle on the same seeded inputs.
ECDSA verdicts are "parity unpinned" by reference vectors: the oracle itself is
cross-checked against oracle/pyref.py and OpenSSL in tests/test_oracle_secp.py."""
import numpy as np
import pytest

from conftest import oracle_ecdsa_batch
from rusty_kaspa_b200 import workload as W

pytestmark = pytest.mark.gpu


def test_ecdsa_parity_mixed_batch(gpu_ctx, oracle):
    pk, msg, sig, kind = W.ecdsa_triples(12000, seed=2, n_keys=1024, n_nonces=1024, frac_bitflip=0.1, frac_adversarial=0.15)
    got = gpu_ctx.verify_ecdsa_batch(pk, msg, sig)
    exp = oracle_ecdsa_batch(oracle, pk, msg, sig)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first at {bad[:5]}: got {got[bad[:5]]} exp {exp[bad[:5]]} kind {kind[bad[:5]]}"
    assert (got[kind == 0] == 1).all()
    assert (got[kind != 0] != 1).all()
    assert set(np.unique(got)) == {0, 1, 2, 3}  # every verdict class is exercised


@pytest.mark.parametrize("n", [0, 1, 33, 129])
def test_ecdsa_ragged_sizes(gpu_ctx, oracle, n):
    pk, msg, sig, kind = W.ecdsa_triples(max(n, 1), seed=n + 11, n_keys=16, n_nonces=16, frac_bitflip=0.2, frac_adversarial=0.2)
    pk, msg, sig = pk[:n], msg[:n], sig[:n]
    got = gpu_ctx.verify_ecdsa_batch(pk, msg, sig, n=n)
    if n:
        assert (got == oracle_ecdsa_batch(oracle, pk, msg, sig)).all()
