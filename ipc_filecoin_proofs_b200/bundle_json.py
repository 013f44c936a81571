"""Wire format of the reference's proof bundles: JSON as `serde_json` renders
`UnifiedProofBundle` / `EventProofBundle` (reference src/proofs/common/bundle.rs:10-45, events/bundle.rs:5-30,
storage/bundle.rs:5-14), built from the engine's POD results. Host-side rendering only — no device work.

Field by field:
  * epochs / indices / actor ids: JSON numbers;
  * CIDs held as `String` (`child_block_cid`, `message_cid`, `parent_tipset_cids`, `parent_state_root`, `actor_state_cid`,
    `storage_root`): `Cid::to_string()` = multibase `b` + lower-case RFC 4648 base32 without padding (`bafy2bzace…`)
    (events/generator.rs:289, storage/generator.rs:170-174);
  * `topics`, `data`, `slot`, `value`: `"0x" + hex::encode(..)`, lower case (events/generator.rs:279-281,
    storage/generator.rs:175-176);
  * `ProofBlock.data`: standard base64 with padding (common/bundle.rs:20-26);
  * `ProofBlock.cid` is a `cid::Cid`, not a String: cid 0.11's `Serialize` hands the CID *bytes* to the serializer
    (`serialize_newtype_struct` → `serialize_bytes`), which serde_json writes as an array of numbers. [UPSTREAM] behaviour
    restated from the published crate; the reference has no fixture that pins it ("parity unpinned").
    `from_json` also accepts the `{"/": "bafy…"}` and plain-string spellings other producers use.
Key order is the struct field order and `dumps` uses serde_json's compact separators, so equal bundles give equal bytes.
"""
import base64
import json

import numpy as np

from . import _abi as A

_B32 = "abcdefghijklmnopqrstuvwxyz234567"
_B32_REV = {c: i for i, c in enumerate(_B32)}


def cid_to_string(cid):
    """38 raw CID bytes → `Cid::to_string()` of a CIDv1: multibase base32 lower, no padding."""
    raw = bytes(cid)
    bits = int.from_bytes(raw, "big")
    nbits = 8 * len(raw)
    pad = (-nbits) % 5
    bits <<= pad
    n = (nbits + pad) // 5
    return "b" + "".join(_B32[(bits >> (5 * (n - 1 - i))) & 31] for i in range(n))


def cid_from_string(s):
    """Inverse of cid_to_string (base32 multibase only — what the reference prints)."""
    if not s or s[0] != "b":
        raise ValueError("only multibase base32 ('b…') CIDs are supported")
    bits = 0
    for ch in s[1:]:
        bits = (bits << 5) | _B32_REV[ch]
    nbits = 5 * (len(s) - 1)
    nbytes = nbits // 8
    extra = nbits - 8 * nbytes
    if bits & ((1 << extra) - 1):
        raise ValueError("non-zero base32 padding bits")
    return (bits >> extra).to_bytes(nbytes, "big")


def _hex(b):
    return "0x" + bytes(b).hex()


def _unhex(s):
    if not s.startswith("0x"):
        raise ValueError("hex fields start with 0x")
    return bytes.fromhex(s[2:])


def proof_blocks(witness):
    """Vec<ProofBlock> in the witness' (Cid) order."""
    return [{"cid": list(bytes(witness.cids[i])), "data": base64.b64encode(witness.block(i)).decode()} for i in range(witness.n_blocks)]


def event_proofs(ts, result):
    """Vec<EventProof> (events/bundle.rs:14-23) of one EventResultPy for tipset pair `ts`."""
    parents = [cid_to_string(c) for c in np.asarray(ts.parent_cids, dtype=np.uint8).reshape(-1, A.CID_LEN)]
    child = cid_to_string(ts.child_cid)
    out = []
    for p in result.proofs:
        out.append({
            "parent_epoch": int(ts.parent_epoch), "child_epoch": int(ts.child_epoch), "parent_tipset_cids": list(parents),
            "child_block_cid": child, "message_cid": cid_to_string(p.message_cid),
            "exec_index": int(p.exec_index), "event_index": int(p.event_index),
            "event_data": {"emitter": int(p.emitter), "topics": [_hex(t) for t in p.topics], "data": _hex(p.data)},
        })
    return out


def storage_proofs(ts, result):
    """Vec<StorageProof> (storage/bundle.rs:5-14) of one StorageResultPy."""
    child = cid_to_string(ts.child_cid)
    psr = cid_to_string(ts.parent_state_root)
    return [{"child_epoch": int(ts.child_epoch), "child_block_cid": child, "parent_state_root": psr, "actor_id": int(p.actor_id),
             "actor_state_cid": cid_to_string(p.actor_state_cid), "storage_root": cid_to_string(p.storage_root),
             "slot": _hex(p.slot), "value": _hex(p.value)} for p in result.proofs]


def event_bundle(ts, result):
    """EventProofBundle {proofs, blocks} (events/bundle.rs:26-30)."""
    return {"proofs": event_proofs(ts, result), "blocks": proof_blocks(result.witness)}


def unified_bundle(ts, bundle):
    """UnifiedProofBundle {storage_proofs, event_proofs, blocks} (common/bundle.rs:37-45) of a BundlePy."""
    sp = storage_proofs(ts, bundle.storage) if bundle.storage is not None else []
    ep = [p for r in bundle.events for p in event_proofs(ts, r)]
    return {"storage_proofs": sp, "event_proofs": ep, "blocks": proof_blocks(bundle.witness)}


def dumps(obj):
    """serde_json::to_string: compact separators, struct field order."""
    return json.dumps(obj, separators=(",", ":"), ensure_ascii=False)


# ------------------------------------------------------------------------------------------ reading bundles back
def _cid_field(v):
    if isinstance(v, list):
        return bytes(v)
    if isinstance(v, dict) and "/" in v:
        return cid_from_string(v["/"])
    if isinstance(v, str):
        return cid_from_string(v)
    raise ValueError("unrecognised CID spelling")


def witness_from_blocks(blocks):
    """Vec<ProofBlock> JSON → WitnessPy (blocks kept in the order given)."""
    cids = np.zeros((len(blocks), A.CID_LEN), dtype=np.uint8)
    offs = np.zeros(len(blocks), dtype=np.uint64)
    lens = np.zeros(len(blocks), dtype=np.uint32)
    blob = bytearray()
    for i, b in enumerate(blocks):
        c = _cid_field(b["cid"])
        if len(c) != A.CID_LEN:
            raise ValueError("only 38-byte CIDv1 (dag-cbor, blake2b-256) CIDs travel through the C ABI")
        data = base64.b64decode(b["data"], validate=True)
        cids[i] = np.frombuffer(c, dtype=np.uint8)
        offs[i] = len(blob)
        lens[i] = len(data)
        blob += data
    return A.WitnessPy(cids, offs, lens, np.frombuffer(bytes(blob) + bytes(16), dtype=np.uint8))


def event_proofs_from_json(items):
    """Vec<EventProof> JSON → list of EventProofPy (tipset fields are returned separately by `tipset_fields`)."""
    return [A.EventProofPy(int(p["exec_index"]), int(p["event_index"]), int(p["event_data"]["emitter"]),
                           [_unhex(t) for t in p["event_data"]["topics"]], _unhex(p["event_data"]["data"]), cid_from_string(p["message_cid"]))
            for p in items]


def storage_proofs_from_json(items):
    return [A.StorageProofPy(int(p["actor_id"]), cid_from_string(p["actor_state_cid"]), cid_from_string(p["storage_root"]), _unhex(p["slot"]),
                             _unhex(p["value"]), True, 32) for p in items]


def loads(text):
    return json.loads(text)
