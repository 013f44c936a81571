"""Loads tests/golden/config1.npz (made by tests/golden/make_golden.py with the independent Python oracle)."""
import os
from types import SimpleNamespace

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config1.npz")


def load():
    z = np.load(PATH)
    ts = SimpleNamespace(
        cids=z["cids"], offsets=z["offsets"], lengths=z["lengths"], blob=z["blob"], n_blocks=len(z["lengths"]),
        parent_epoch=int(z["parent_epoch"]), child_epoch=int(z["child_epoch"]), n_parents=len(z["parent_cids"]),
        parent_cids=z["parent_cids"], parent_txmeta_cids=z["parent_txmeta_cids"], child_cid=z["child_cid"], receipts_root=z["receipts_root"],
        parent_state_root=z["parent_state_root"], n_receipts=len(z["has_events_root"]), events_roots=z["events_roots"],
        has_events_root=z["has_events_root"], event_signature=str(z["event_signature"]), topic1=str(z["topic1"]), actor_filter=None)
    s = SimpleNamespace(
        cids=z["s_cids"], offsets=z["s_offsets"], lengths=z["s_lengths"], blob=z["s_blob"], n_blocks=len(z["s_lengths"]),
        parent_epoch=0, child_epoch=1, n_parents=len(z["s_parent_cids"]), parent_cids=z["s_parent_cids"],
        parent_txmeta_cids=z["s_parent_txmeta_cids"], child_cid=z["s_child_cid"], receipts_root=z["s_receipts_root"],
        parent_state_root=z["s_parent_state_root"], n_receipts=0, events_roots=np.zeros((0, 38), np.uint8), has_events_root=np.zeros(0, np.uint8))
    return z, ts, s


def check_event_result(z, r):
    assert r.matching.tolist() == z["exp_matching"].tolist()
    assert np.array_equal(r.witness.cids, z["exp_witness"])
    assert r.n_exec == len(z["exp_exec"])
    ep = z["exp_proofs"]
    assert len(r.proofs) == len(ep)
    to = do = 0
    for k, p in enumerate(r.proofs):
        assert (p.exec_index, p.event_index, p.emitter) == tuple(int(x) for x in ep[k])
        tb = b"".join(p.topics)
        assert tb == z["exp_proof_topics"][to:to + len(tb)].tobytes()
        to += len(tb)
        assert p.data == z["exp_proof_data"][do:do + len(p.data)].tobytes()
        do += len(p.data)
        assert p.message_cid == z["exp_proof_msg"][k].tobytes()
    # every witness block hashes to its CID
    import hashlib
    for i in range(r.witness.n_blocks):
        assert hashlib.blake2b(r.witness.block(i), digest_size=32).digest() == r.witness.cids[i][6:].tobytes()


def check_storage_result(z, res):
    n = len(z["s_actor"])
    assert len(res.proofs) == n
    wo = 0
    for k, p in enumerate(res.proofs):
        assert p.found == bool(z["s_found"][k])
        assert p.value == z["s_value"][k].tobytes()
        assert p.actor_state_cid == z["s_state_cid"][k].tobytes()
        assert p.storage_root == z["s_storage_root"][k].tobytes()
        wl = int(z["s_witness_len"][k])
        exp_w = z["s_witness"][wo:wo + wl]
        wo += wl
        got_w = res.witness.cids[res.spec_witness[k]]
        assert np.array_equal(got_w, exp_w)


# ------------------------------------------------------------------ Keccak-256 known answers held by the reference tree itself
REF_KECCAK_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_keccak_vectors.json")


def check_reference_keccak_vectors(keccak256):
    """Evaluates every vector of tests/golden/reference_keccak_vectors.json (extracted from the reference's vendored forge-std by
    tests/golden/make_reference_keccak_vectors.py) with the given `keccak256(bytes) -> 32 bytes`. Returns the number of hash calls."""
    import json
    doc = json.load(open(REF_KECCAK_PATH))
    calls = [0]

    def k(m):
        calls[0] += 1
        d = bytes(keccak256(bytes(m)))
        assert len(d) == 32
        return d

    def addr(s):
        return bytes.fromhex(s[2:])

    def create(deployer20, nonce):   # keccak256(rlp([deployer, nonce]))[12:], nonce < 0x80 (single RLP byte, 0 → 0x80)
        assert 0 <= nonce < 0x80
        return k(bytes([0xc0 + 22, 0x80 + 20]) + deployer20 + (bytes([nonce]) if nonce else b"\x80"))[12:]

    for v in doc["vectors"]:
        kind, src = v["kind"], v["source"]
        if kind == "digest":
            assert k(bytes.fromhex(v["message_hex"])).hex() == v["expect"][2:], src
        elif kind == "low20":
            assert k(v["message_ascii"].encode())[12:] == addr(v["expect"]), src
        elif kind == "prefix4":
            assert k(v["message_ascii"].encode())[:4].hex() == v["expect"][2:], src
        elif kind == "create2":   # keccak256(0xff ‖ deployer ‖ salt ‖ keccak256(initcode))[12:]
            init_hash = k(bytes.fromhex(v["initcode_preimage_hex"]))
            assert k(b"\xff" + addr(v["deployer"]) + bytes.fromhex(v["salt_hex"]) + init_hash)[12:] == addr(v["expect"]), src
        elif kind == "create_chain":
            a = addr(v["deployer"])
            for n in v["nonces"]:
                a = create(a, n)
            assert a == addr(v["expect"]), src
        elif kind == "eip55":   # EIP-55: hex digit i is upper case iff nibble i of keccak256(lower-case hex ascii) >= 8
            lower = v["address"][2:].lower()
            d = k(lower.encode()).hex()
            assert "".join(c.upper() if c in "abcdef" and int(d[i], 16) >= 8 else c for i, c in enumerate(lower)) == v["address"][2:], src
        else:
            raise AssertionError(kind)
    assert len(doc["vectors"]) >= 17
    return calls[0]
