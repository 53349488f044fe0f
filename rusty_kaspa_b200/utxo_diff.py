"""Host-side mirror of the reference's UTXO diff algebra: `UtxoDiff` with `with_diff`, `diff_from`,
`add_transaction`, reversal — consensus/core/src/utxo/utxo_diff.rs:15-262 and utxo_collection.rs.  This is bookkeeping the node keeps
on the host (one diff per chain block, a few hundred entries); the GPU table consumes diffs through GpuUtxoSet.apply_diff /
add_transactions.  Pinned by the reference's own rule table (tests/golden/utxo_diff_rules.json, tests/test_utxo_diff.py).

An outpoint is (txid: bytes, index: int); an entry is a dict with amount, spk_version, script, block_daa_score, is_coinbase."""


class UtxoAlgebraError(Exception):
    """kinds as in consensus/core/src/utxo/utxo_error.rs:7-25; equality of errors is by kind and outpoint"""

    def __init__(self, kind, outpoint=None):
        super().__init__("%s %r" % (kind, outpoint))
        self.kind, self.outpoint = kind, outpoint


def _has(coll, outpoint, daa_score):  # contains_with_daa_score
    e = coll.get(outpoint)
    return e is not None and e["block_daa_score"] == daa_score


class UtxoDiff:
    def __init__(self, add=None, remove=None):
        self.add = dict(add or {})
        self.remove = dict(remove or {})

    def __eq__(self, other):
        return isinstance(other, UtxoDiff) and self.add == other.add and self.remove == other.remove

    def __repr__(self):
        return "UtxoDiff(add=%r, remove=%r)" % (self.add, self.remove)

    def clone(self):
        return UtxoDiff(self.add, self.remove)

    def to_reversed(self):
        return UtxoDiff(self.remove, self.add)

    # ---- with_diff (utxo_diff.rs:76-116): self then other, applied to one base set
    def with_diff(self, other):
        c = self.clone()
        c.with_diff_in_place(other)
        return c

    def with_diff_in_place(self, other):
        for o in sorted(other.remove.keys() & self.remove.keys()):
            if not _has(self.add, o, other.remove[o]["block_daa_score"]):
                raise UtxoAlgebraError("DuplicateRemovePoint", o)
        for o in sorted(other.add.keys() & self.add.keys()):
            if not _has(other.remove, o, self.add[o]["block_daa_score"]):
                raise UtxoAlgebraError("DuplicateAddPoint", o)
        # what other removes: cancels an addition of ours made at the same DAA score, otherwise it is a removal from the base
        cancelled = [o for o, e in other.remove.items() if _has(self.add, o, e["block_daa_score"])]
        for o, e in other.remove.items():
            if o not in cancelled:
                self.remove[o] = e
        for o in cancelled:
            del self.add[o]
        # what other adds: cancels a removal of ours of the same DAA score, otherwise it is an addition
        cancelled = [o for o, e in other.add.items() if _has(self.remove, o, e["block_daa_score"])]
        for o, e in other.add.items():
            if o not in cancelled:
                self.add[o] = e
        for o in cancelled:
            del self.remove[o]

    # ---- diff_from (utxo_diff.rs:118-225): the diff that turns self into other, both taken from one base set
    def diff_from(self, other):
        for o in sorted(self.remove.keys() & other.add.keys()):
            t, x = self.remove[o], other.add[o]
            if not (x["block_daa_score"] != t["block_daa_score"]
                    and (_has(self.add, o, x["block_daa_score"]) or _has(other.remove, o, t["block_daa_score"]))):
                raise UtxoAlgebraError("DiffIntersectionPoint", o)
        for o in sorted(self.add.keys() & other.remove.keys()):
            t, x = self.add[o], other.remove[o]
            if not (x["block_daa_score"] != t["block_daa_score"]
                    and (_has(self.remove, o, x["block_daa_score"]) or _has(other.add, o, t["block_daa_score"]))):
                raise UtxoAlgebraError("DiffIntersectionPoint", o)
        for o in sorted(self.remove.keys() & other.remove.keys()):
            if self.remove[o]["block_daa_score"] != other.remove[o]["block_daa_score"]:
                raise UtxoAlgebraError("DiffIntersectionPoint", o)
        res = UtxoDiff()
        in_both = {}
        for o, e in self.add.items():  # our additions the other side does not have must be undone
            if _has(other.add, o, e["block_daa_score"]):
                in_both[o] = e
            else:
                res.remove[o] = e
        if bool(in_both.keys() & self.remove.keys()) != bool(in_both.keys() & other.remove.keys()):
            raise UtxoAlgebraError("General")
        for o, e in other.remove.items():  # removals only the other side made
            if not _has(self.remove, o, e["block_daa_score"]):
                res.remove[o] = e
        for o, e in self.remove.items():  # removals only we made must be restored
            if not _has(other.remove, o, e["block_daa_score"]):
                res.add[o] = e
        for o, e in other.add.items():  # additions only the other side made
            if not _has(self.add, o, e["block_daa_score"]):
                res.add[o] = e
        return res

    # ---- add_transaction (utxo_diff.rs:227-261)
    def add_transaction(self, tx, entries, tx_id, block_daa_score, is_coinbase=False):
        """tx: tx dict (txbatch layout); entries: the populated entry of every input; tx_id: bytes"""
        for i, e in zip(tx["inputs"], entries):
            o = (i["txid"], i["index"])
            if _has(self.add, o, e["block_daa_score"]):
                del self.add[o]
            elif o not in self.remove:
                self.remove[o] = e
            else:
                raise UtxoAlgebraError("DoubleRemoveCall", o)
        for k, out in enumerate(tx["outputs"]):
            o = (tx_id, k)
            e = {"amount": out["value"], "spk_version": out["spk_version"], "script": out["script"], "block_daa_score": block_daa_score, "is_coinbase": is_coinbase}
            if _has(self.remove, o, block_daa_score):
                del self.remove[o]
            elif o not in self.add:
                self.add[o] = e
            else:
                raise UtxoAlgebraError("DoubleAddCall", o)

    # ---- handing a diff to the GPU table (DbUtxoSetStore::write_diff_batch order: removals, then additions)
    def apply_to(self, gpu_utxo_set):
        import numpy as np
        from .simgen import entries_to_arrays
        key = lambda o: o[0] + int(o[1]).to_bytes(4, "little")
        rk = np.frombuffer(b"".join(key(o) for o in self.remove), dtype=np.uint8).reshape(-1, 36) if self.remove else None
        ak = np.frombuffer(b"".join(key(o) for o in self.add), dtype=np.uint8).reshape(-1, 36) if self.add else None
        ae, ab = entries_to_arrays(list(self.add.values())) if self.add else (None, None)
        return gpu_utxo_set.apply_diff(rem_keys36=rk, add_keys36=ak, add_entries=ae, add_bytes=ab)
