"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle (bit-exact)."""
import hashlib

import numpy as np
import pytest

from ipc_filecoin_proofs_b200 import _abi as A
from tests.util import ShuffledTipset, assert_event_results_equal, assert_witness_equal, spec_of

pytestmark = pytest.mark.gpu

LENS = [0, 1, 2, 31, 32, 55, 56, 63, 64, 65, 111, 112, 127, 128, 129, 135, 136, 137, 255, 256, 257, 271, 272, 273, 1027, 1028, 1029, 4096, 5000]


def _msgs():
    rng = np.random.default_rng(1)
    return [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in LENS]


def test_blake2b_batch(api):
    msgs = _msgs()
    got = api.blake2b256_batch(msgs)
    assert got == [hashlib.blake2b(m, digest_size=32).digest() for m in msgs]
    assert got[0].hex() == "0e5751c026e543b2e8ab2eb06099daa1d1e5df47778f7787faab45cdf12fe3a8"


def test_sha256_batch(api):
    msgs = _msgs()
    assert api.sha256_batch(msgs) == [hashlib.sha256(m).digest() for m in msgs]


def test_keccak_batch(api, oracle_mod):
    msgs = _msgs() + [b"Transfer(address,address,uint256)", b"NewTopDownMessage(bytes32,uint256)"]
    got = api.keccak256_batch(msgs)
    assert got == [oracle_mod.keccak256(m) for m in msgs]
    assert got[0].hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert got[-2].hex() == "ddf252ad1be2c89b69c2b068fc378daa952ba7f163c4a11628f55a4df523b3ef"


def test_mapping_slots(api, oracle_mod):
    keys = [bytes([i]) * 32 for i in range(5)] + [b"calib-subnet-1" + bytes(18)]
    idx = [0, 1, 2, 2**40, 2**64 - 1, 0]
    got = api.compute_mapping_slots(keys, idx)
    assert got == [oracle_mod.compute_mapping_slot(k, i) for k, i in zip(keys, idx)]
    assert api.calculate_storage_slot("calib-subnet-1", 0) == got[-1]


def test_store_get_has_verify(api, ts1):
    st = api.BlockStore.from_tipset(ts1, verify_cids=True)
    for i in (0, 5, ts1.n_blocks - 1):
        assert st.get(ts1.cids[i]) == ts1.block(i)
        assert st.has(ts1.cids[i])
    assert st.get(np.zeros(38, dtype=np.uint8)) is None
    assert not st.has(bytes([1, 0x71, 0xa0, 0xe4, 2, 0x20]) + bytes(32))
    # corrupt one byte of one block → CID mismatch at that block
    blob = ts1.blob.copy()
    victim = 17
    blob[int(ts1.offsets[victim]) + 3] ^= 0x40
    with pytest.raises(A.IpcfpError) as ei:
        api.BlockStore(ts1.cids, ts1.offsets, ts1.lengths, blob, verify_cids=True)
    assert ei.value.status == A.ERR_CID_MISMATCH and ei.value.index == victim


@pytest.mark.parametrize("cfg", [1, 2])
def test_event_proof_parity(api, oracle_mod, synth_mod, cfg):
    ts = synth_mod.Tipset(synth_mod.config_params(cfg))
    exp = oracle_mod.Store.from_tipset(ts).generate_event_proof(ts, spec_of(ts))
    got = api.BlockStore.from_tipset(ts, verify_cids=True).generate_event_proof(ts, spec_of(ts))
    assert_event_results_equal(got, exp)
    assert got.matching.tolist() == ts.selected.tolist()
    assert all(oracle_mod.verify_event_proofs(got.witness, ts, got, spec_of(ts)))


@pytest.mark.parametrize("kw", [
    dict(n_receipts=300, events_per_receipt=40, match_ppm=100000),                       # two-level events AMTs (bw 5) / three-level (bw 3)
    dict(n_receipts=500, events_per_receipt=3, null_root_permille=200, match_ppm=200000),  # receipts without events root
    dict(n_receipts=257, events_per_receipt=8, bw3_permille=1000, match_ppm=50000, has_actor_filter=0),
    dict(n_receipts=1000, events_per_receipt=8, case_a_permille=500, malformed_permille=100, match_ppm=30000),
    dict(n_receipts=1, events_per_receipt=1, match_ppm=1000000, dup_msgs=0, n_parents=1),
    dict(n_receipts=9, events_per_receipt=8, match_ppm=0, n_parents=3, dup_msgs=2),
    dict(n_receipts=700, events_per_receipt=300, match_ppm=20000, n_parents=1),             # bw-3 AMTs of height 2
])
def test_event_proof_shapes(api, oracle_mod, synth_mod, kw):
    ts = synth_mod.Tipset(synth_mod.default_params(seed=99, **kw))
    exp = oracle_mod.Store.from_tipset(ts).generate_event_proof(ts, spec_of(ts))
    got = api.BlockStore.from_tipset(ts).generate_event_proof(ts, spec_of(ts))
    assert_event_results_equal(got, exp)


def test_event_proof_block_order_and_alignment(api, oracle_mod, ts2):
    exp = oracle_mod.Store.from_tipset(ts2).generate_event_proof(ts2, spec_of(ts2))
    for misalign in (False, True):
        sh = ShuffledTipset(ts2, seed=3, misalign=misalign)
        got = api.BlockStore.from_tipset(sh, verify_cids=True).generate_event_proof(sh, spec_of(sh))
        assert_event_results_equal(got, exp)


def test_skip_tx_flag(api, oracle_mod, ts2):
    exp = oracle_mod.Store.from_tipset(ts2).generate_event_proof(ts2, spec_of(ts2), flags=A.SCAN_SKIP_TX_AMTS)
    got = api.BlockStore.from_tipset(ts2).generate_event_proof(ts2, spec_of(ts2), flags=A.SCAN_SKIP_TX_AMTS)
    assert_event_results_equal(got, exp)


def test_storage_slots_parity(api, oracle_mod, ts3_small):
    ts = ts3_small
    n = int(ts.params.hamt_entries)
    rng = np.random.default_rng(5)
    ks = rng.integers(0, n, 900).tolist() + [n]
    keys = [ts.storage_entry(k)[0] for k in ks] + [ts.storage_absent_key(k) for k in range(100)]
    slots = api.compute_mapping_slots(keys, [0] * len(keys))
    slots_np = np.frombuffer(b"".join(slots), dtype=np.uint8)
    exp = oracle_mod.Store.from_tipset(ts).read_storage_slots(ts.storage_root, slots_np)
    got = api.BlockStore.from_tipset(ts, verify_cids=True).read_storage_slots(ts.storage_root, slots_np)
    assert np.array_equal(got.found, exp.found) and np.array_equal(got.raw_len, exp.raw_len) and np.array_equal(got.values, exp.values)
    assert_witness_equal(got.witness, exp.witness)
    assert got.found[:901].all() and not got.found[901:].any()
    for i, k in enumerate(ks[:50]):
        v = ts.storage_entry(k)[1]
        assert bytes(got.values[i][32 - len(v):]) == v


def test_storage_proofs_parity(api, oracle_mod, ts3_small):
    ts = ts3_small
    n = int(ts.params.hamt_entries)
    keys = [ts.storage_entry(k)[0] for k in (0, 1, 2, 77, n)] + [ts.storage_absent_key(1)]
    slots = api.compute_mapping_slots(keys, [0] * len(keys))
    specs = [(actor, s) for actor in (1001, 1002, 1003, 1004, 1005, 1006) for s in slots]
    exp = oracle_mod.Store.from_tipset(ts).generate_storage_proofs(ts, specs)
    got = api.BlockStore.from_tipset(ts).generate_storage_proofs(ts, specs)
    assert [vars(p) for p in got.proofs] == [vars(p) for p in exp.proofs]
    assert_witness_equal(got.witness, exp.witness)
    assert got.spec_witness == exp.spec_witness
    assert all(oracle_mod.verify_storage_proofs(got.witness, ts, got))


def test_bundle_parity(api, oracle_mod, ts3_small):
    ts = ts3_small
    slot = api.calculate_storage_slot("calib-subnet-1", 0)
    sspecs = [(1001, slot), (1003, slot)]
    especs = [A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter), A.make_event_spec(ts.event_signature, "calib-subnet-2", None)]
    exp = oracle_mod.Store.from_tipset(ts).generate_proof_bundle(ts, sspecs, especs)
    got = api.BlockStore.from_tipset(ts).generate_proof_bundle(ts, sspecs, especs)
    assert [vars(p) for p in got.storage.proofs] == [vars(p) for p in exp.storage.proofs]
    for g, e in zip(got.events, exp.events):
        assert_event_results_equal(g, e)
    assert_witness_equal(got.witness, exp.witness)
