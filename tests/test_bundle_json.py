"""JSON wire format of the proof bundles (reference common/bundle.rs, events/bundle.rs, storage/bundle.rs).
CPU only: the POD results come from the oracle — they have exactly the layout the engine returns through the C ABI."""
import base64
import json
from types import SimpleNamespace

import numpy as np
import pytest

from ipc_filecoin_proofs_b200 import _abi as A
from ipc_filecoin_proofs_b200 import bundle_json as J
from tests.util import spec_of


def test_cid_strings():
    c = bytes.fromhex("0171a0e40220") + bytes(range(32))
    s = J.cid_to_string(c)
    assert s.startswith("bafy2bzace") and len(s) == 62          # what Cid::to_string() prints for Filecoin chain CIDs (reference main.rs output)
    assert s[1:] == base64.b32encode(c).decode().lower().rstrip("=")
    assert J.cid_from_string(s) == c
    for n in (1, 4, 5, 36, 38, 40):                              # every padding case of base32
        raw = bytes((7 * i + n) & 0xff for i in range(n))
        assert J.cid_from_string(J.cid_to_string(raw)) == raw


def test_event_bundle_json_round_trip(oracle_mod, ts1):
    spec = spec_of(ts1)
    res = oracle_mod.Store.from_tipset(ts1).generate_event_proof(ts1, spec)
    assert res.proofs
    text = J.dumps(J.event_bundle(ts1, res))
    assert ": " not in text and ", " not in text               # serde_json::to_string is compact
    doc = json.loads(text)
    assert list(doc) == ["proofs", "blocks"]
    p0 = doc["proofs"][0]
    assert list(p0) == ["parent_epoch", "child_epoch", "parent_tipset_cids", "child_block_cid", "message_cid", "exec_index", "event_index",
                        "event_data"]                                                       # struct field order, events/bundle.rs:14-23
    assert list(p0["event_data"]) == ["emitter", "topics", "data"]
    assert p0["parent_epoch"] == int(ts1.parent_epoch) and p0["child_epoch"] == int(ts1.child_epoch)
    assert len(p0["parent_tipset_cids"]) == int(ts1.n_parents) and all(c.startswith("bafy2bzace") for c in p0["parent_tipset_cids"])
    assert all(t.startswith("0x") and len(t) == 66 and t == t.lower() for t in p0["event_data"]["topics"])
    b0 = doc["blocks"][0]
    assert list(b0) == ["cid", "data"] and b0["cid"][:6] == [1, 0x71, 0xa0, 0xe4, 0x02, 0x20] and len(b0["cid"]) == 38
    # back to POD and through the restated verifier (events/verifier.rs)
    w = J.witness_from_blocks(doc["blocks"])
    assert np.array_equal(w.cids, res.witness.cids) and w.blocks() == res.witness.blocks()
    proofs = J.event_proofs_from_json(doc["proofs"])
    assert [p.key() for p in proofs] == [p.key() for p in res.proofs]
    raw, blob = A.pack_event_proofs(proofs)
    back = SimpleNamespace(proofs=proofs, raw_proofs=raw, data_blob=blob)
    assert all(oracle_mod.verify_event_proofs(w, ts1, back, spec))
    # a corrupted block every replay needs (the child header) must not verify
    k = [i for i, x in enumerate(doc["blocks"]) if bytes(x["cid"]) == bytes(ts1.child_cid)]
    assert len(k) == 1
    doc["blocks"][k[0]]["data"] = base64.b64encode(b"\x80").decode()
    try:
        ok = any(oracle_mod.verify_event_proofs(J.witness_from_blocks(doc["blocks"]), ts1, back, spec))
    except A.IpcfpError:
        ok = False
    assert not ok


def test_unified_bundle_json_round_trip(oracle_mod, ts3_small):
    ts = ts3_small
    slot = oracle_mod.compute_mapping_slot((b"calib-subnet-1" + bytes(32))[:32], 0)
    sspecs = [(1001, slot), (1003, slot)]
    especs = [A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)]
    b = oracle_mod.Store.from_tipset(ts).generate_proof_bundle(ts, sspecs, especs)
    doc = json.loads(J.dumps(J.unified_bundle(ts, b)))
    assert list(doc) == ["storage_proofs", "event_proofs", "blocks"]            # common/bundle.rs:37-45
    s0 = doc["storage_proofs"][0]
    assert list(s0) == ["child_epoch", "child_block_cid", "parent_state_root", "actor_id", "actor_state_cid", "storage_root", "slot", "value"]
    assert s0["slot"] == "0x" + bytes(slot).hex() and len(s0["value"]) == 66
    assert len(doc["blocks"]) == b.witness.n_blocks
    # storage proofs verify from the JSON alone (storage/verifier.rs)
    w = J.witness_from_blocks(doc["blocks"])
    sp = J.storage_proofs_from_json(doc["storage_proofs"])
    back = SimpleNamespace(proofs=sp, raw_proofs=A.pack_storage_proofs(sp))
    assert all(oracle_mod.verify_storage_proofs(w, ts, back))
    ep = J.event_proofs_from_json(doc["event_proofs"])
    assert [p.key() for p in ep] == [p.key() for r in b.events for p in r.proofs]
    raw, blob = A.pack_event_proofs(ep)
    assert all(oracle_mod.verify_event_proofs(w, ts, SimpleNamespace(proofs=ep, raw_proofs=raw, data_blob=blob), especs[0]))
    # the other CID spellings consumers use are accepted on input
    alt = [{"cid": {"/": J.cid_to_string(bytes(x["cid"]))}, "data": x["data"]} for x in doc["blocks"][:3]]
    assert np.array_equal(J.witness_from_blocks(alt).cids, w.cids[:3])


def test_c_abi_json_equals_python_rendering(oracle_mod, ts3_small, ts1):
    """ipcfp_bundle_to_json / ipcfp_event_result_to_json (csrc/bundle_json.cpp: host-side rendering behind the C ABI, no device) give
    byte for byte what bundle_json.py gives — on PODs produced by the oracle, which have exactly the layout the engine returns."""
    import ctypes as C
    from ipc_filecoin_proofs_b200 import api
    ts = ts3_small
    slot = oracle_mod.compute_mapping_slot((b"calib-subnet-1" + bytes(32))[:32], 0)
    sspecs = [(1001, slot), (1003, slot)]
    especs = [A.make_event_spec(ts.event_signature, ts.topic1, ts.actor_filter)]
    OL = oracle_mod.lib()
    ost = oracle_mod.Store.from_tipset(ts)
    d, keep = A.make_tipset_desc(ts)
    sarr = A.make_storage_specs(sspecs)
    earr = (A.EventSpec * len(especs))(*especs)
    out = C.POINTER(A.BundleC)()
    assert OL.oracle_generate_proof_bundle(ost._h, C.byref(d), sarr, len(sspecs), earr, len(especs), C.byref(out)) == 0
    try:
        text = api.bundle_to_json(out, ts)
        want = J.dumps(J.unified_bundle(ts, A.bundle_from_c(out.contents)))
        assert text == want
        assert json.loads(text)["storage_proofs"][0]["slot"] == "0x" + bytes(slot).hex()
    finally:
        OL.oracle_bundle_free(out)
    # EventProofBundle of one result
    spec = spec_of(ts1)
    ost1 = oracle_mod.Store.from_tipset(ts1)
    d1, keep1 = A.make_tipset_desc(ts1)
    eo = C.POINTER(A.EventResultC)()
    assert OL.oracle_generate_event_proof(ost1._h, C.byref(d1), C.byref(spec), 0, 1, C.byref(eo)) == 0
    try:
        text = api.event_result_to_json(eo, ts1)
        assert text == J.dumps(J.event_bundle(ts1, A.event_result_from_c(eo.contents)))
        # a by-reference witness (IPCFP_WITNESS_BY_REFERENCE: blob == NULL) has no bytes to render: refused, not dereferenced
        byref = A.EventResultC.from_buffer_copy(eo.contents)
        byref.witness.blob = None
        byref.witness.blob_size = 0
        with pytest.raises(A.IpcfpError) as ei:
            api.event_result_to_json(C.pointer(byref), ts1)
        assert ei.value.status == A.ERR_INVALID_ARG
    finally:
        OL.oracle_event_result_free(eo)
