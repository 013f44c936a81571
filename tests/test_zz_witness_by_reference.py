"""IPCFP_WITNESS_BY_REFERENCE (include/ipcfp.h): the witness without block bytes — CIDs / lengths in `Cid` order as always, offsets
into the blob the store was created from. Everything else of the result is unchanged, and the blocks the offsets name are byte for
byte the blocks the default mode copies. (Last file of the suite on purpose: an opt-in mode must not stand in front of the others.)"""
import numpy as np
import pytest

from ipc_filecoin_proofs_b200 import _abi as A
from tests.util import ShuffledTipset, assert_event_results_equal, spec_of

pytestmark = pytest.mark.gpu


def _check(api, oracle_mod, ts):
    exp = oracle_mod.Store.from_tipset(ts).generate_event_proof(ts, spec_of(ts))
    store = api.BlockStore.from_tipset(ts, verify_cids=True)
    full = store.generate_event_proof(ts, spec_of(ts))
    ref = store.generate_event_proof(ts, spec_of(ts), flags=A.WITNESS_BY_REFERENCE)
    assert_event_results_equal(full, exp)
    assert ref.matching.tolist() == full.matching.tolist()
    assert [p.key() for p in ref.proofs] == [p.key() for p in full.proofs] and ref.n_exec == full.n_exec
    assert np.array_equal(ref.witness.cids, full.witness.cids) and np.array_equal(ref.witness.lengths, full.witness.lengths)
    assert len(ref.witness.blob) == 0
    blob = np.asarray(ts.blob, dtype=np.uint8)
    for i in range(full.witness.n_blocks):
        o, n = int(ref.witness.offsets[i]), int(ref.witness.lengths[i])
        assert bytes(blob[o:o + n]) == full.witness.block(i), i


@pytest.mark.parametrize("cfg", [1, 2])
def test_by_reference_equals_copied_witness(api, oracle_mod, synth_mod, cfg):
    _check(api, oracle_mod, synth_mod.Tipset(synth_mod.config_params(cfg)))


def test_by_reference_shuffled_misaligned_store(api, oracle_mod, ts2):
    _check(api, oracle_mod, ShuffledTipset(ts2, seed=5, misalign=True))


def test_by_reference_general_walk(api, oracle_mod, synth_mod, monkeypatch):
    monkeypatch.setenv("IPCFP_BFS_GENERAL", "1")
    _check(api, oracle_mod, synth_mod.Tipset(synth_mod.default_params(seed=7, n_receipts=700, events_per_receipt=5, match_ppm=50000, n_parents=3, dup_msgs=4)))
