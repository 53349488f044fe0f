"""Pruning-point UTXO-set import on the GPU, pinned by data the reference wrote: a chain block of its simpa DAG fixture plays the
new pruning point; the UTXO set in its past leaves one GPU table as the store's RocksDB rows and enters a fresh table chunk by chunk
(append_imported_pruning_point_utxos, consensus/src/consensus/mod.rs:1070-1083); the finalized multiset must equal the utxoCommitment in THAT header
and the block's own transactions must validate against the imported set (import_pruning_point_utxo_set, processor.rs:1126-1172)."""
import numpy as np
import pytest

from rusty_kaspa_b200 import Params
from rusty_kaspa_b200.txbatch import build_batch

pytestmark = pytest.mark.gpu


def _replay_chain_until(gpu_ctx, fixture, want_txs):
    """the virtual chain of the fixture replayed on the GPU (as tests/test_gpu_replay.py does) up to the LAST chain block whose own body holds at
    least `want_txs` non-coinbase transactions; returns (fixture params, block record, table holding the UTXO set of that block's past)"""
    from golden_util import simpa_dag_replay_plan
    from rusty_kaspa_b200 import GpuUtxoSet, TransactionValidator
    from rusty_kaspa_b200.validator import FLAGS_FULL, FLAGS_SKIP_SCRIPT_CHECKS
    fx, by, order, sp, ordered_mergeset, chain = simpa_dag_replay_plan(fixture)
    stop = max(i for i, b in enumerate(chain) if i > 0 and len(by[b]["txs"]) - 1 >= want_txs)
    tv = TransactionValidator(gpu_ctx, Params(coinbase_maturity=fx["coinbase_maturity"], storage_mass_parameter=fx["storage_mass_parameter"]))
    us = GpuUtxoSet(gpu_ctx, 1 << 16)
    for b in chain[1:stop + 1]:
        pov, s = by[b]["daa_score"], sp(b)
        us.add_transactions(build_batch([by[s]["txs"][0]]), np.ones(1, dtype=np.uint8), pov)
        for k, mb in enumerate(ordered_mergeset(b)):
            txs = by[mb]["txs"][1:]
            if not txs:
                continue
            batch = build_batch(txs)
            res = tv.validate_transactions_in_parallel(us, batch, pov, FLAGS_SKIP_SCRIPT_CHECKS if k == 0 else FLAGS_FULL)
            us.add_transactions(batch, (res["status"] == 0).astype(np.uint8), pov)
    return fx, by[chain[stop]], us


def test_pruning_point_utxo_set_import_reproduces_the_reference_commitment(gpu_ctx):
    from rusty_kaspa_b200 import GpuUtxoSet, MuHash, store_rows
    from rusty_kaspa_b200.pruning_import import PruningPointImport, ImportedMultisetHashMismatch
    fx, pp, src = _replay_chain_until(gpu_ctx, "simpa_goref_1060.json.gz", 2)
    commitment = bytes.fromhex(pp["utxo_commitment"])
    assert MuHash.of_utxo_set(gpu_ctx, src).finalize() == commitment  # the source side is where the reference says it is
    # ---- syncer side: the set leaves the table as store rows
    keys, ent, arena = src.export()
    n = len(keys)
    assert n == src.count() and n > 300
    assert len({k.tobytes() for k in keys}) == n
    key_rows, key_off, value_rows, value_off = store_rows.encode_rows(keys, ent, arena)
    # ---- syncee side: chunks of rows into a fresh table + running multiset
    prm = Params(coinbase_maturity=fx["coinbase_maturity"], storage_mass_parameter=fx["storage_mass_parameter"])
    dst = GpuUtxoSet(gpu_ctx, 1 << 14)
    imp = PruningPointImport(gpu_ctx, dst, prm)
    chunk = 97
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        ko, vo = np.asarray(key_off[a:b + 1]), np.asarray(value_off[a:b + 1])
        imp.append_rows(bytes(key_rows[int(ko[0]):int(ko[-1])]), ko - ko[0], bytes(value_rows[int(vo[0]):int(vo[-1])]), vo - vo[0])
    assert imp.n_imported == n and dst.count() == n and dst.digest() == src.digest()
    batch = build_batch(pp["txs"])
    res = imp.import_pruning_point_utxo_set(commitment, batch, pp["daa_score"])
    assert int(res["status"][0]) == 12 and (res["status"][1:] == 0).all() and len(res) >= 3
    # the same transactions against the SOURCE table: identical verdicts and fees
    ref = imp.tv.validate_transactions_in_parallel(src, batch, pp["daa_score"])
    assert (ref["status"] == res["status"]).all() and (ref["fee"] == res["fee"]).all()
    # ---- a chunk altered in transit: ImportedMultisetHashMismatch (processor.rs:1134-1139)
    bad = GpuUtxoSet(gpu_ctx, 1 << 14)
    imp2 = PruningPointImport(gpu_ctx, bad, prm)
    ent2 = ent.copy()
    ent2["amount"][n // 2] += 1
    for a in range(0, n, 500):
        imp2.append_imported_pruning_point_utxos(keys[a:a + 500], ent2[a:a + 500], arena)
    with pytest.raises(ImportedMultisetHashMismatch):
        imp2.import_pruning_point_utxo_set(commitment, batch, pp["daa_score"])
    # ---- and the export / import pair is lossless for every script class the table stores (inline and overflow scripts)
    rng = np.random.default_rng(11)
    from rusty_kaspa_b200.txbatch import ENTRY_DTYPE
    m = 4000
    k2 = rng.integers(0, 256, size=(m, 36), dtype=np.uint8)
    k2[:, 33:] = 0                                   # small indices: the rows' key trimming is exercised
    k2[::7, 32] = 0
    lens = rng.choice([0, 1, 34, 35, 37, 68, 69, 150, 400], size=m)
    e2 = np.zeros(m, dtype=ENTRY_DTYPE)
    e2["amount"] = rng.integers(1, 1 << 50, size=m)
    e2["block_daa_score"] = rng.integers(0, 1 << 40, size=m)
    e2["spk_version"] = rng.integers(0, 3, size=m)
    e2["is_coinbase"] = rng.integers(0, 2, size=m)
    e2["script_len"] = lens
    e2["script_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]])
    a2 = rng.integers(0, 256, size=int(lens.sum()) + 8, dtype=np.uint8)
    t1, t2 = GpuUtxoSet(gpu_ctx, 1 << 14), GpuUtxoSet(gpu_ctx, 1 << 13)
    t1.apply_diff(add_keys36=k2, add_entries=e2, add_bytes=a2)
    kk, ee, aa = t1.export()
    imp3 = PruningPointImport(gpu_ctx, t2, prm)
    for a in range(0, m, 1000):
        imp3.append_imported_pruning_point_utxos(kk[a:a + 1000], ee[a:a + 1000], aa)
    assert t2.count() == t1.count() == m and t2.digest() == t1.digest()
    assert imp3.multiset().finalize() == MuHash.of_utxo_set(gpu_ctx, t1).finalize()
    for t in (src, dst, bad, t1, t2):
        t.close()
