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


def test_pass2_one_match_per_thread_variant(api, oracle_mod, synth_mod, monkeypatch):
    """k_pass2 runs one matching receipt per warp up to 16 384 matches and one per thread above; IPCFP_PASS2_PER_THREAD forces the latter
    so that both launch shapes stay covered at test sizes (same per-item code: pass2_item)."""
    monkeypatch.setenv("IPCFP_PASS2_PER_THREAD", "1")
    for ts in (synth_mod.Tipset(synth_mod.config_params(2)),
               synth_mod.Tipset(synth_mod.default_params(seed=11, n_receipts=3000, events_per_receipt=6, match_ppm=200000, n_parents=2, dup_msgs=3))):
        exp = oracle_mod.Store.from_tipset(ts).generate_event_proof(ts, spec_of(ts))
        got = api.BlockStore.from_tipset(ts, verify_cids=True).generate_event_proof(ts, spec_of(ts))
        assert_event_results_equal(got, exp)


def test_public_filecoin_constants_on_the_gpu(api):
    """The chain's own constants (see tests/test_oracle_cpu.py::test_public_filecoin_constants_pin_the_encodings) through the GPU's
    Blake2b-256: empty v0 AMT → empty TxMeta, builtin-actors' EMPTY_ARR_CID, the empty HAMT node."""
    from ipc_filecoin_proofs_b200 import bundle_json as J
    from tests.test_oracle_cpu import (EMPTY_AMT_V0, EMPTY_AMT_V3, EMPTY_HAMT_NODE, FILECOIN_EMPTY_ARR, FILECOIN_EMPTY_HAMT, FILECOIN_EMPTY_TXMETA)
    pre = bytes([0x01, 0x71, 0xa0, 0xe4, 0x02, 0x20])
    d = api.blake2b256_batch([EMPTY_AMT_V0, EMPTY_AMT_V3, EMPTY_HAMT_NODE])
    cid_v0 = pre + bytes(d[0])
    txmeta = bytes([0x82]) + (bytes([0xd8, 0x2a, 0x58, 0x27, 0x00]) + cid_v0) * 2
    d2 = api.blake2b256_batch([txmeta])
    assert J.cid_to_string(pre + bytes(d2[0])) == FILECOIN_EMPTY_TXMETA
    assert J.cid_to_string(pre + bytes(d[1])) == FILECOIN_EMPTY_ARR
    assert J.cid_to_string(pre + bytes(d[2])) == FILECOIN_EMPTY_HAMT


def test_public_solidity_storage_layout_vectors_on_the_gpu(api):
    """k_mapping_slots / k_hash_batch<keccak> against the public Solidity storage-layout vectors (tests/test_oracle_cpu.py)."""
    from tests.test_oracle_cpu import SOLIDITY_ARRAY_VECTORS, SOLIDITY_SLOT_VECTORS
    got = api.compute_mapping_slots([k for k, _, _ in SOLIDITY_SLOT_VECTORS], [i for _, i, _ in SOLIDITY_SLOT_VECTORS])
    assert [bytes(g).hex() for g in got] == [w for _, _, w in SOLIDITY_SLOT_VECTORS]
    got = api.keccak256_batch([m for m, _ in SOLIDITY_ARRAY_VECTORS])
    assert [bytes(g).hex() for g in got] == [w for _, w in SOLIDITY_ARRAY_VECTORS]


def test_keccak_vectors_held_by_the_reference_tree_on_the_gpu(api):
    """k_hash_batch<keccak> against the Keccak-256 constants the reference's own tree holds (vendored forge-std; fixture
    tests/golden/reference_keccak_vectors.json, see tests/test_oracle_cpu.py::test_keccak_vectors_held_by_the_reference_tree)."""
    from tests.golden_util import check_reference_keccak_vectors
    assert check_reference_keccak_vectors(lambda m: api.keccak256_batch([m])[0]) >= 20


def test_json_bundle_verified_through_the_c_abi_alone(api, oracle_mod, synth_mod):
    """The flow a non-Rust host has: EventProofBundle as JSON text → ipcfp_bundle_from_json → witness store with every block
    Blake2b-checked → ipcfp_verify_event_proofs, with the tipset fields exactly as the parser recovered them from the proofs."""
    import ctypes as C
    from ipc_filecoin_proofs_b200 import bundle_json as J
    ts = synth_mod.Tipset(synth_mod.config_params(2))
    spec = spec_of(ts)
    r = api.BlockStore.from_tipset(ts, verify_cids=True).generate_event_proof(ts, spec)
    assert len(r.proofs) > 0
    pb = api.ParsedBundle(J.dumps(J.event_bundle(ts, r)))
    L = api.lib()
    w = pb.c.witness
    store = C.c_void_p()
    assert L.ipcfp_store_create(w.cids, w.offsets, w.lengths, w.blob, w.blob_size, w.n_blocks, 0, A.STORE_VERIFY_CIDS, C.byref(store)) == 0, L.ipcfp_last_error()
    try:
        n = int(pb.c.n_event_proofs)
        assert n == len(r.proofs)
        res = np.zeros(n, dtype=np.uint8)
        st = L.ipcfp_verify_event_proofs(store, C.byref(pb.c.tipset), pb.c.event_proofs, n, pb.c.data_blob, pb.c.data_blob_size, C.addressof(spec),
                                         res.ctypes.data)
        assert st == 0, L.ipcfp_last_error()
        assert res.all()
        assert [bool(x) for x in res] == oracle_mod.verify_event_proofs(r.witness, ts, r, spec)
    finally:
        L.ipcfp_store_destroy(store)
        pb.close()
