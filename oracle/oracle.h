/* oracle.h — CPU restatement of the reference's hot path (TEST INFRASTRUCTURE).
 *
 * PARITY UNPINNED BY THE REFERENCE: consensus-shipyard/ipc-filecoin-proofs ships no tests,
 * fixtures or golden vectors, cannot be compiled here (no Rust toolchain, crates
 * not vendored) and delegates all arithmetic to un-vendored crates (fvm_ipld_amt 0.7.4,
 * fvm_ipld_hamt 0.10.4, fvm_ipld_encoding 0.5.3, serde_ipld_dagcbor 0.6, fvm_shared 4.7, cid 0.11,
 * multihash-codetable 0.1.4, sha3 0.10 — Cargo.toml:11-22,38). This oracle restates the reference's
 * orchestration line by line and the crates' published formats; it is pinned instead by
 *   (1) known-answer hash vectors + hashlib,
 *   (2) an independent Python (cbor2 + hashlib) builder/scanner (oracle/pyoracle.py),
 *   (3) the restated verifiers (every witness must verify; dropping a block must fail),
 *   (4) three constants of the public Filecoin chain (empty TxMeta over v0 AMTs, builtin-actors' EMPTY_ARR_CID, the empty HAMT
 *       node): tests/test_oracle_cpu.py::test_public_filecoin_constants_pin_the_encodings — the basic encodings (DAG-CBOR tuples
 *       and links, AMT v0/v3 and HAMT node layouts, Blake2b-256 CIDs) are checked against the real network, not only against
 *       ourselves.
 *   (5) the only known answers the reference TREE holds for anything on the path: Keccak-256 constants of its vendored forge-std
 *       (tests/golden/reference_keccak_vectors.json, extracted by tests/golden/make_reference_keccak_vectors.py) — they pin
 *       oracle_keccak256 / oracle_compute_mapping_slot's hash, nothing else.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The product library never links or calls it.
 *
 * Result structs are the POD types of include/ipcfp.h so tests compare field by field.
 */
#ifndef IPCFP_ORACLE_H
#define IPCFP_ORACLE_H
#include "../include/ipcfp.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_store oracle_store;

/* Borrows all four arrays for the lifetime of the store (MemoryBlockstore whose get() clones). */
oracle_store* oracle_store_create(const uint8_t* cids, const uint64_t* offsets, const uint32_t* lengths,
                                  const uint8_t* blob, uint64_t n_blocks);
void oracle_store_destroy(oracle_store* s);
/* first block whose blake2b-256 differs from its CID digest, or UINT64_MAX */
uint64_t oracle_store_verify_cids(const oracle_store* s, uint32_t threads);

const char* oracle_last_error(void);
uint64_t oracle_last_error_index(void);

/* generate_event_proof (events/generator.rs:60-107). threads = 1 mirrors the reference
 * (single-threaded); threads > 1 parallelises pass 1 over receipts. */
ipcfp_status oracle_generate_event_proof(const oracle_store* s, const ipcfp_tipset_desc* t, const ipcfp_event_spec* spec,
                                         uint32_t flags, uint32_t threads, ipcfp_event_result** out);
/* scan receipts [lo,hi) only (multi-GPU shard semantics of ipcfp_generate_event_proof_shard) */
ipcfp_status oracle_generate_event_proof_shard(const oracle_store* s, const ipcfp_tipset_desc* t,
                                               const ipcfp_event_spec* spec, uint64_t lo, uint64_t hi, uint32_t world,
                                               uint32_t rank, uint32_t flags, uint32_t threads, ipcfp_event_result** out);
void oracle_event_result_free(ipcfp_event_result* r);

ipcfp_status oracle_read_storage_slots(const oracle_store* s, const uint8_t root[38], const uint8_t* slots, uint64_t k,
                                       ipcfp_slot_result** out);
void oracle_slot_result_free(ipcfp_slot_result* r);

ipcfp_status oracle_generate_storage_proofs(const oracle_store* s, const ipcfp_tipset_desc* t,
                                            const ipcfp_storage_spec* specs, uint64_t n, ipcfp_storage_result** out);
void oracle_storage_result_free(ipcfp_storage_result* r);

ipcfp_status oracle_generate_proof_bundle(const oracle_store* s, const ipcfp_tipset_desc* t, const ipcfp_storage_spec* ss,
                                          uint64_t ns, const ipcfp_event_spec* es, uint64_t ne, ipcfp_bundle** out);
void oracle_bundle_free(ipcfp_bundle* b);

/* verify_event_proof (events/verifier.rs:51-290) over a witness; trust policy = accept all;
 * filter_spec may be NULL (create_event_filter, verifier.rs:28-40). results[n_proofs] ∈ {0,1}. */
ipcfp_status oracle_verify_event_proofs(const ipcfp_witness* w, const ipcfp_tipset_desc* t, const ipcfp_event_proof* proofs,
                                        uint64_t n_proofs, const uint8_t* data_blob, const ipcfp_event_spec* filter_spec,
                                        uint8_t* results);
/* verify_storage_proof (storage/verifier.rs:24-63) */
ipcfp_status oracle_verify_storage_proofs(const ipcfp_witness* w, const ipcfp_tipset_desc* t,
                                          const ipcfp_storage_proof* proofs, uint64_t n_proofs, uint8_t* results);

/* unit helpers */
/* TEST HOOK (differential fuzzing, tests/host_fuzz): decode ONE StampedEvent item starting at p (events/generator.rs:218 via
 * fvm_shared StampedEvent) and run extract_evm_log on it (common/evm.rs:13-59). Returns IPCFP_OK or IPCFP_ERR_DECODE.
 * consumed = bytes of the item; some = log is Some; topics_out receives min(ntopics, cap/32) topics; data_len the full length. */
ipcfp_status oracle_decode_event(const uint8_t* p, uint64_t n, uint64_t* consumed, uint64_t* emitter, uint32_t* some, uint32_t* ntopics,
                                 uint8_t* topics_out, uint64_t topics_cap, uint8_t* data_out, uint64_t data_cap, uint64_t* data_len);

/* TEST HOOK: what pass 1 does with ONE events-AMT root block: Amt::<StampedEvent>::load (v3) + for_each + extract_evm_log
 * (events/generator.rs:215-236) over a store that holds only this block (a child link therefore fails as a missing block).
 * Per visited event (up to cap): AMT index, emitter, Some/None, number of topics, data length. */
ipcfp_status oracle_scan_events_block(const uint8_t* block, uint64_t n, uint64_t* n_events, uint64_t* idx, uint64_t* emitter, uint8_t* some,
                                      uint32_t* ntopics, uint64_t* dlen, uint64_t cap);

/* TEST HOOK: one node of an Amtv0<Receipt> (non-root: [bmap, [links], [receipts]]) at the given height, whole-node decode as
 * fvm_ipld_amt does it (events/generator.rs:196,249). Out: number of links / values; per value (up to cap) whether it has an
 * events root and the 38 CID bytes. */
ipcfp_status oracle_decode_receipts_node(const uint8_t* p, uint64_t n, uint32_t height, uint32_t* n_links, uint32_t* n_vals, uint8_t* has_root,
                                         uint8_t* roots38, uint64_t cap);

/* TEST HOOK: one HAMT node (fvm_ipld_hamt v3 layout), whole-node decode, then the pointer of slot `idx` for `key`
 * (storage/decode.rs:79-96, common/decode.rs:29-39). vkind 0 = ActorState values, 1 = Vec<u8> values (CBOR array of u8).
 * kind: 0 None, 1 value (out = ActorState.state CID, 38 bytes / the u8 elements), 2 link (out = the 38 CID bytes). */
ipcfp_status oracle_hamt_node_lookup(const uint8_t* p, uint64_t n, int vkind, uint32_t idx, const uint8_t* key, uint32_t keylen, int32_t* kind,
                                     uint8_t* out, uint64_t out_cap, uint64_t* out_len);

/* TEST HOOK: the concatenated ("raw") message list of a tipset — every BLS then SECP message AMT of every parent block,
 * in order, BEFORE the first-seen dedup (events/utils.rs:48-94) — as 38-byte CIDs. */
ipcfp_status oracle_message_list(const oracle_store* s, const ipcfp_tipset_desc* t, uint8_t* out38, uint64_t cap, uint64_t* n);

void oracle_keccak256(const uint8_t* in, uint64_t len, uint8_t out[32]);
void oracle_blake2b256(const uint8_t* in, uint64_t len, uint8_t out[32]);
void oracle_sha256(const uint8_t* in, uint64_t len, uint8_t out[32]);
void oracle_compute_mapping_slot(const uint8_t key32[32], uint64_t slot_index, uint8_t out[32]);
/* sorts n 38-byte CIDs in `Cid` Ord order in place, removing duplicates; returns new count */
uint64_t oracle_sort_unique_cids(uint8_t* cids, uint64_t n);

#ifdef __cplusplus
}
#endif
#endif
