"""Helpers shared by the parity tests."""
import numpy as np

from ipc_filecoin_proofs_b200 import _abi as A


def spec_of(ts):
    return A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)


def assert_event_results_equal(got, exp, check_witness_bytes=True):
    assert got.matching.tolist() == exp.matching.tolist()
    assert got.n_exec == exp.n_exec
    assert [p.key() for p in got.proofs] == [p.key() for p in exp.proofs]
    assert got.witness.n_blocks == exp.witness.n_blocks
    assert np.array_equal(got.witness.cids, exp.witness.cids)
    if check_witness_bytes:
        assert np.array_equal(got.witness.lengths, exp.witness.lengths)
        assert got.witness.blocks() == exp.witness.blocks()
    assert np.array_equal(got.data_blob, exp.data_blob)


def assert_witness_equal(a, b):
    assert np.array_equal(a.cids, b.cids)
    assert np.array_equal(a.lengths, b.lengths)
    assert a.blocks() == b.blocks()


class ShuffledTipset:
    """Same tipset with the flat block arrays permuted (the engine must not depend on block order)."""

    def __init__(self, ts, seed=7, misalign=False):
        rng = np.random.default_rng(seed)
        perm = rng.permutation(ts.n_blocks)
        self._ts = ts
        lens = ts.lengths[perm]
        offs = np.zeros(ts.n_blocks, dtype=np.uint64)
        pos = 0
        chunks = []
        for k, i in enumerate(perm):
            pad = int(rng.integers(0, 7)) if misalign else (16 - pos % 16) % 16
            chunks.append(bytes(pad))
            pos += pad
            offs[k] = pos
            b = ts.block(int(i))
            chunks.append(b)
            pos += len(b)
        self.blob = np.frombuffer(b"".join(chunks) + bytes(16), dtype=np.uint8)
        self.cids = ts.cids[perm].copy()
        self.offsets = offs
        self.lengths = lens.copy()
        self.n_blocks = ts.n_blocks

    def __getattr__(self, name):
        return getattr(self._ts, name)


class EditedTipset:
    """Tipset view with replaced arrays (for fault injection)."""

    def __init__(self, ts, **over):
        self._ts = ts
        for k in ("cids", "offsets", "lengths", "blob", "n_blocks"):
            setattr(self, k, over.get(k, getattr(ts, k)))
        for k, v in over.items():
            setattr(self, k, v)

    def __getattr__(self, name):
        return getattr(self._ts, name)


def dict_of(ts):
    """cid -> bytes of any tipset-like object, from its flat arrays (first occurrence wins)."""
    d = {}
    for i in range(int(ts.n_blocks)):
        o = int(ts.offsets[i])
        d.setdefault(bytes(ts.cids[i]), bytes(ts.blob[o:o + int(ts.lengths[i])]))
    return d
