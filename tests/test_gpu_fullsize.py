"""BASELINE.json full-size batch (1 Mi Schnorr triples, configs[1]) through size-independent properties:
ground truth by construction, batch-split invariance, permutation equivariance, bitmap consistency, and an
oracle spot check on a random sample (the full oracle pass would take minutes of CPU)."""
import numpy as np
import pytest

from conftest import oracle_schnorr_batch
from rusty_kaspa_b200 import workload as W

pytestmark = pytest.mark.gpu
N = 1 << 20


@pytest.fixture(scope="module")
def triples():
    return W.schnorr_triples(N, seed=0x6B61737061)


def test_full_batch_properties(gpu_ctx, oracle, triples):
    pk, msg, sig, kind = triples
    st = gpu_ctx.verify_schnorr_batch(pk, msg, sig)
    # 1. ground truth by construction: untouched items valid, corrupted / adversarial ones never valid
    assert (st[kind == 0] == 1).all() and not (st[kind != 0] == 1).any()
    assert set(np.unique(st)) <= {0, 1, 2}
    # 2. results do not depend on how the batch is split (determinism, independent items)
    cut = 333_333
    a = gpu_ctx.verify_schnorr_batch(pk[:cut].copy(), msg[:cut].copy(), sig[:cut].copy())
    b = gpu_ctx.verify_schnorr_batch(pk[cut:].copy(), msg[cut:].copy(), sig[cut:].copy())
    assert (np.concatenate([a, b]) == st).all()
    # 3. permutation equivariance
    perm = np.random.default_rng(1).permutation(N)
    stp = gpu_ctx.verify_schnorr_batch(pk[perm].copy(), msg[perm].copy(), sig[perm].copy())
    assert (stp == st[perm]).all()
    # 4. bitmap = packed (status == valid); popcount = number of valid items
    bm = gpu_ctx.status_to_bitmap(st)
    assert (bm == np.packbits((st == 1).astype(np.uint8), bitorder="little")).all()
    assert int(np.unpackbits(bm).sum()) == int((kind == 0).sum())
    # 5. oracle spot check: every non-valid item (about 2 %) plus a random sample of valid ones
    idx = np.concatenate([np.nonzero(kind != 0)[0][:6000], np.random.default_rng(2).choice(N, 6000, replace=False)])
    exp = oracle_schnorr_batch(oracle, pk[idx].copy(), msg[idx].copy(), sig[idx].copy())
    assert (st[idx] == exp).all()
