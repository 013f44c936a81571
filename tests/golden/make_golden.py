#!/usr/bin/env python
"""Generates tests/golden/config1.npz: the BASELINE.json configs[0] tipset (64 receipts x 8 events,
single topic_0 filter) as flat block arrays + the expected results computed by the INDEPENDENT
Python oracle (oracle/pyoracle.py: cbor2 + hashlib), plus storage-path vectors on a small HAMT.

The reference itself ships no golden vectors and cannot be built here; this
fixture pins the C++ oracle and the CUDA engine against a second implementation.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from oracle import pyoracle as P  # noqa: E402


def main():
    ts = synth.Tipset(synth.config_params(1))
    store = ts.as_dict()
    r = P.generate_event_proof(store, ts, ts.event_signature, ts.topic1, ts.actor_filter)
    out = dict(
        cids=ts.cids.copy(), offsets=ts.offsets.copy(), lengths=ts.lengths.copy(), blob=ts.blob.copy(),
        parent_epoch=ts.parent_epoch, child_epoch=ts.child_epoch, parent_cids=ts.parent_cids.copy(),
        parent_txmeta_cids=ts.parent_txmeta_cids.copy(), child_cid=ts.child_cid.copy(), receipts_root=ts.receipts_root.copy(),
        parent_state_root=ts.parent_state_root.copy(), events_roots=ts.events_roots.copy(), has_events_root=ts.has_events_root.copy(),
        event_signature=ts.event_signature, topic1=ts.topic1,
        exp_matching=np.array(r["matching"], dtype=np.uint64),
        exp_witness=np.frombuffer(b"".join(r["witness"]), dtype=np.uint8).reshape(-1, 38),
        exp_exec=np.frombuffer(b"".join(r["exec_order"]), dtype=np.uint8).reshape(-1, 38),
        exp_proofs=np.array([(i, j, em) for (i, j, em, t, d, m) in r["proofs"]], dtype=np.uint64),
        exp_proof_topics=np.frombuffer(b"".join(b"".join(t) for (_, _, _, t, _, _) in r["proofs"]), dtype=np.uint8),
        exp_proof_data=np.frombuffer(b"".join(d for (_, _, _, _, d, _) in r["proofs"]), dtype=np.uint8),
        exp_proof_msg=np.frombuffer(b"".join(m for (_, _, _, _, _, m) in r["proofs"]), dtype=np.uint8).reshape(-1, 38),
    )
    # storage path on a 2000-entry HAMT
    ts3 = synth.Tipset(synth.config_params(3, hamt_entries=2000, n_receipts=8))
    st3 = ts3.as_dict()
    specs, exp = [], []
    for actor in (1001, 1002, 1003, 1004, 1005, 1006):
        for k in (0, 1, 2, 1999, 2000, -1):
            key = ts3.storage_absent_key(7) if k < 0 else ts3.storage_entry(k)[0]
            slot = P.compute_mapping_slot(key, 0)
            a = P.generate_storage_proof(st3, ts3, actor, slot)
            specs.append((actor, slot))
            exp.append((a["found"], a["value"], a["actor_state_cid"], a["storage_root"], b"".join(a["witness"])))
    out.update(
        s_cids=ts3.cids.copy(), s_offsets=ts3.offsets.copy(), s_lengths=ts3.lengths.copy(), s_blob=ts3.blob.copy(),
        s_child_cid=ts3.child_cid.copy(), s_parent_state_root=ts3.parent_state_root.copy(), s_receipts_root=ts3.receipts_root.copy(),
        s_parent_cids=ts3.parent_cids.copy(), s_parent_txmeta_cids=ts3.parent_txmeta_cids.copy(),
        s_actor=np.array([a for a, _ in specs], dtype=np.uint64),
        s_slot=np.frombuffer(b"".join(s for _, s in specs), dtype=np.uint8).reshape(-1, 32),
        s_found=np.array([e[0] for e in exp], dtype=np.uint8),
        s_value=np.frombuffer(b"".join(e[1] for e in exp), dtype=np.uint8).reshape(-1, 32),
        s_state_cid=np.frombuffer(b"".join(e[2] for e in exp), dtype=np.uint8).reshape(-1, 38),
        s_storage_root=np.frombuffer(b"".join(e[3] for e in exp), dtype=np.uint8).reshape(-1, 38),
        s_witness_len=np.array([len(e[4]) // 38 for e in exp], dtype=np.uint32),
        s_witness=np.frombuffer(b"".join(e[4] for e in exp), dtype=np.uint8).reshape(-1, 38),
    )
    # hash known-answer vectors (hashlib for blake2b/sha256; keccak from the published vectors + pyoracle)
    msgs = [bytes(range(256))[:n] * 1 for n in (0, 1, 55, 56, 64, 127, 128, 129, 135, 136, 137, 255)]
    import hashlib
    out.update(
        kat_msgs=np.frombuffer(b"".join(msgs), dtype=np.uint8), kat_lens=np.array([len(m) for m in msgs], dtype=np.uint32),
        kat_blake2b=np.frombuffer(b"".join(hashlib.blake2b(m, digest_size=32).digest() for m in msgs), dtype=np.uint8).reshape(-1, 32),
        kat_sha256=np.frombuffer(b"".join(hashlib.sha256(m).digest() for m in msgs), dtype=np.uint8).reshape(-1, 32),
        kat_keccak=np.frombuffer(b"".join(P.keccak256(m) for m in msgs), dtype=np.uint8).reshape(-1, 32),
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
