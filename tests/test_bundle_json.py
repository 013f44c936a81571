"""JSON wire format of the proof bundles (reference common/bundle.rs, events/bundle.rs, storage/bundle.rs).
CPU only: the POD results come from the oracle — they have exactly the layout the engine returns through the C ABI."""
import ctypes as C
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


def test_c_abi_json_parser_round_trip(oracle_mod, ts3_small, ts1):
    """ipcfp_bundle_from_json (csrc/bundle_parse.cpp): JSON → the PODs the batched verifiers take + the witness block arrays. The parsed
    PODs equal what the Python parser + packers produce from the same text, byte for byte, and the restated verifiers accept them."""
    from ipc_filecoin_proofs_b200 import api
    # --- UnifiedProofBundle
    ts = ts3_small
    slot = oracle_mod.compute_mapping_slot((b"calib-subnet-1" + bytes(32))[:32], 0)
    ost = oracle_mod.Store.from_tipset(ts)
    b = ost.generate_proof_bundle(ts, [(1001, slot), (1003, slot)], [spec_of(ts)])
    text = J.dumps(J.unified_bundle(ts, b))
    pb = api.ParsedBundle(text)
    doc = J.loads(text)
    w_py = J.witness_from_blocks(doc["blocks"])
    w_c = pb.witness
    assert np.array_equal(w_c.cids, w_py.cids) and np.array_equal(w_c.lengths, w_py.lengths)
    assert [w_c.block(i) for i in range(w_c.n_blocks)] == [w_py.block(i) for i in range(w_py.n_blocks)]
    assert all(int(o) % 16 == 0 for o in w_c.offsets)
    ep = J.event_proofs_from_json(doc["event_proofs"])
    raw_c, blob_c = pb.event_proofs_raw
    raw_py, blob_py = A.pack_event_proofs(ep)
    assert raw_c.tobytes() == raw_py.tobytes() and blob_c.tobytes() == blob_py.tobytes()[:len(blob_c)]
    sp = J.storage_proofs_from_json(doc["storage_proofs"])
    assert pb.storage_proofs_raw.tobytes() == A.pack_storage_proofs(sp).tobytes()
    f = pb.tipset_fields()
    assert f["child_epoch"] == int(ts.child_epoch) and f["child_cid"] == bytes(ts.child_cid)
    if doc["event_proofs"]:   # parent epoch / parent tipset CIDs travel with the event proofs only (events/bundle.rs:14-23)
        assert f["parent_epoch"] == int(ts.parent_epoch) and f["parent_cids"] == bytes(np.asarray(ts.parent_cids, dtype=np.uint8).reshape(-1))
    assert f["parent_state_root"] == bytes(ts.parent_state_root)
    # closed loop: the restated verifiers on what the C parser produced
    assert all(oracle_mod.verify_event_proofs(w_c, ts, SimpleNamespace(proofs=ep, raw_proofs=raw_c, data_blob=blob_c), spec_of(ts)))
    assert all(oracle_mod.verify_storage_proofs(w_c, ts, SimpleNamespace(proofs=sp, raw_proofs=pb.storage_proofs_raw)))
    pb.close()
    # --- EventProofBundle, with the other CID spellings of ProofBlock.cid and an unknown field (serde ignores it)
    r1 = oracle_mod.Store.from_tipset(ts1).generate_event_proof(ts1, spec_of(ts1))
    d1 = J.event_bundle(ts1, r1)
    for k, blk in enumerate(d1["blocks"]):
        if k % 3 == 1:
            blk["cid"] = {"/": J.cid_to_string(bytes(blk["cid"]))}
        elif k % 3 == 2:
            blk["cid"] = J.cid_to_string(bytes(blk["cid"]))
    d1["note"] = {"ignored": [1, 2.5e3, None, True, "é\\"]}
    pb1 = api.ParsedBundle(json.dumps(d1))
    assert np.array_equal(pb1.witness.cids, r1.witness.cids) and pb1.witness.blocks() == r1.witness.blocks()
    raw1, blob1 = pb1.event_proofs_raw
    assert raw1.tobytes() == A.pack_event_proofs(r1.proofs)[0].tobytes()
    assert pb1.c.n_storage_proofs == 0 and not pb1.c.tipset.child_parent_state_root
    f1 = pb1.tipset_fields()
    assert (f1["parent_epoch"], f1["child_epoch"]) == (int(ts1.parent_epoch), int(ts1.child_epoch))
    assert f1["parent_cids"] == bytes(np.asarray(ts1.parent_cids, dtype=np.uint8).reshape(-1)) and f1["child_cid"] == bytes(ts1.child_cid)
    # the flow of tests/test_zz_witness_by_reference.py::test_json_bundle_verified_through_the_c_abi_alone with the restated verifier in
    # the GPU verifier's place: witness, proofs AND tipset descriptor exactly as the C parser recovered them (no TxMeta CIDs, no receipts root)
    n1 = int(pb1.c.n_event_proofs)
    res = np.zeros(n1, dtype=np.uint8)
    sp1 = spec_of(ts1)
    st = oracle_mod.lib().oracle_verify_event_proofs(C.byref(pb1.c.witness), C.byref(pb1.c.tipset), pb1.c.event_proofs, n1, pb1.c.data_blob,
                                                     C.byref(sp1), res.ctypes.data)
    assert st == 0 and n1 == len(r1.proofs) > 0 and res.all()
    pb1.close()


def test_c_abi_json_parser_refuses_malformed_input(oracle_mod, ts1):
    from ipc_filecoin_proofs_b200 import api
    r1 = oracle_mod.Store.from_tipset(ts1).generate_event_proof(ts1, spec_of(ts1))
    good = J.dumps(J.event_bundle(ts1, r1))
    api.ParsedBundle(good).close()

    def status(text):
        try:
            api.ParsedBundle(text).close()
            return A.OK
        except A.IpcfpError as e:
            return e.status

    rng = np.random.default_rng(5)
    for cut in rng.integers(1, len(good) - 1, 200):                       # every truncation is malformed JSON
        assert status(good[:int(cut)]) == A.ERR_INVALID_ARG
    assert status(good + " x") == A.ERR_INVALID_ARG                        # trailing characters
    assert status("[]") == A.ERR_INVALID_ARG and status('{"blocks":[]}') == A.ERR_INVALID_ARG
    doc = json.loads(good)
    assert doc["proofs"], "the fixture has proofs"

    def mutated(fn):
        d = json.loads(good)
        fn(d)
        return json.dumps(d)

    assert status(mutated(lambda d: d["proofs"][0].__setitem__("exec_index", -1))) == A.ERR_INVALID_ARG
    assert status(mutated(lambda d: d["proofs"][0].__setitem__("exec_index", 1.5))) == A.ERR_INVALID_ARG
    assert status(mutated(lambda d: d["proofs"][0].__setitem__("exec_index", 2 ** 64))) == A.ERR_INVALID_ARG
    assert status(mutated(lambda d: d["proofs"][0]["event_data"].__setitem__("data", "0xabc"))) == A.ERR_INVALID_ARG        # odd hex
    assert status(mutated(lambda d: d["proofs"][0]["event_data"].__setitem__("data", "abcd"))) == A.ERR_INVALID_ARG         # no 0x
    assert status(mutated(lambda d: d["proofs"][0].__setitem__("message_cid", "bafy!"))) == A.ERR_INVALID_ARG
    assert status(mutated(lambda d: d["proofs"][0].__setitem__("message_cid", "baeaaa"))) == A.ERR_UNSUPPORTED                 # not a 38-byte CID
    assert status(mutated(lambda d: d["blocks"][0].__setitem__("data", d["blocks"][0]["data"][:-1]))) == A.ERR_INVALID_ARG    # base64 length
    assert status(mutated(lambda d: d["blocks"][0].__setitem__("cid", d["blocks"][0]["cid"][:-1]))) == A.ERR_UNSUPPORTED
    assert status(mutated(lambda d: d["proofs"][0]["event_data"].__setitem__("topics", ["0x00"]))) == A.ERR_UNSUPPORTED       # topic not 32 bytes
    if len(doc["proofs"]) > 1:                                                                                                # proofs of two tipsets in one bundle
        assert status(mutated(lambda d: d["proofs"][1].__setitem__("child_epoch", d["proofs"][1]["child_epoch"] + 1))) == A.ERR_UNSUPPORTED
    assert status(mutated(lambda d: d["proofs"][0].pop("event_index"))) == A.ERR_INVALID_ARG                                  # missing field


def test_c_abi_json_parser_mutation_fuzz_under_sanitizers(oracle_mod, ts3_small, ts1, tmp_path):
    """tests/host_fuzz/fuzz_json.cpp: csrc/bundle_parse.cpp built with AddressSanitizer + UBSan (plain g++), fed mutated copies of an
    EventProofBundle and a UnifiedProofBundle as exact-size heap buffers. Accepted texts must name only bytes inside the returned
    object's own buffers; refused ones must carry a documented status; no over-read of the input, no leak."""
    import os
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    build = os.path.join(root, "tests", "host_fuzz", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "fuzz_json")
    cc = subprocess.run([gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-o", exe,
                         os.path.join(root, "tests", "host_fuzz", "fuzz_json.cpp"), os.path.join(root, "ipc_filecoin_proofs_b200", "csrc", "bundle_parse.cpp"),
                         os.path.join(root, "ipc_filecoin_proofs_b200", "csrc", "bundle_json.cpp")], capture_output=True, text=True)
    if cc.returncode != 0 and "sanitize" in cc.stderr:
        pytest.skip("this g++ has no sanitizer runtime")
    assert cc.returncode == 0, cc.stderr[-2000:]
    r1 = oracle_mod.Store.from_tipset(ts1).generate_event_proof(ts1, spec_of(ts1))
    slot = oracle_mod.compute_mapping_slot((b"calib-subnet-1" + bytes(32))[:32], 0)
    b = oracle_mod.Store.from_tipset(ts3_small).generate_proof_bundle(ts3_small, [(1001, slot), (1003, slot)], [spec_of(ts3_small)])
    seeds = []
    for name, doc in (("ev.json", J.event_bundle(ts1, r1)), ("uni.json", J.unified_bundle(ts3_small, b))):
        p = tmp_path / name
        p.write_text(J.dumps(doc))
        seeds.append(str(p))
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
    out = subprocess.run([exe, "6000", "20260923"] + seeds, capture_output=True, text=True, env=env)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    assert out.stdout.startswith("ok: 12000 mutated bundles parsed"), out.stdout
    accepted = int(out.stdout.split("parsed:")[1].split()[0])
    assert 200 < accepted < 11000, out.stdout   # both outcomes must be exercised
