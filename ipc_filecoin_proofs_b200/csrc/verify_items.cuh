// verify_items.cuh — the per-item device code of the GPU-batched verifiers (verify.cu): what ONE thread does for the tipset, for one
// TxMeta block, for one event proof, for one storage proof. Kept apart from the kernels so that tests/host_fuzz/emu_verify.cu can
// compile the very same code for the host and run it, item by item, against the restated CPU verifiers of the test tree — on intact and on
// adversarial bundles, under AddressSanitizer. The kernels in verify.cu only compute the item index and call these.
#pragma once
#include "engine.cuh"
#include "events_items.cuh"
#include "hashes.cuh"
#include "storage.cuh"

namespace ipcfp {

#define ST_VERIFY 9u

// HeaderLite (common/decode.rs:100-118) fields the verifier needs
struct HeaderFields { uint32_t parents_off, n_parents, psr_off, receipts_off, messages_off; int64_t height; };
__device__ __forceinline__ void header_fields(Rd& r, HeaderFields& h) {
    rd_array_exact(r, 16);
    for (int i = 0; i < 5; i++) rd_skip_any(r);
    h.n_parents = rd_array(r);
    h.parents_off = r.pos;
    for (uint32_t i = 0; i < h.n_parents && !r.err; i++) (void)rd_cid(r);
    rd_skip_any(r);
    h.height = rd_int(r);
    h.psr_off = rd_cid(r);
    h.receipts_off = rd_cid(r);
    h.messages_off = rd_cid(r);
    rd_skip_any(r);
    (void)rd_uint(r);
    rd_skip_any(r);
    (void)rd_uint(r);
    rd_skip_any(r);
    rd_end(r);
}

struct VerifyTipsetArgs {
    StoreView store;
    const uint8_t* parent_cids;   // device, n_parents*38 (from the proof / the caller's tipset)
    const uint8_t* child_cid;     // device, 38
    uint32_t n_parents;
    int64_t parent_epoch, child_epoch;
    // outputs
    uint32_t* consistent;         // verify_header_consistency returned true
    uint32_t* receipts_root_blk;  // block of child_hdr.parent_message_receipts (0xffffffff: not in the witness — an error only if a proof gets that far)
    uint8_t* txmeta_cids;         // n_parents*38: hdr.messages of every parent header
    unsigned long long* err;
};
// once per call, one thread: verify_header_consistency (events/verifier.rs:147-181), the parent headers' `messages` links
// (reconstruct_execution_order, events/utils.rs:16-30) and the TxMeta recompute (utils.rs:64-73)
__device__ __forceinline__ void verify_tipset_item(const VerifyTipsetArgs& a) {
    const StoreView& s = a.store;
    *a.consistent = 0;
    *a.receipts_root_blk = 0xffffffffu;
    int32_t cb = store_lookup(s, a.child_cid);
    if (cb < 0) { report_error(a.err, ST_VERIFY, 0, DC_MISSING, 1); return; }
    uint32_t cl;
    const uint8_t* cp = store_block(s, (uint32_t)cb, cl);
    Rd cr(cp, cl);
    HeaderFields ch;
    header_fields(cr, ch);
    if (cr.err) { report_error(a.err, ST_VERIFY, 0, DC_DECODE, cr.err); return; }
    bool same = ch.n_parents == a.n_parents;
    for (uint32_t k = 0; same && k < a.n_parents; k++) same = cid38_equal(cp + ch.parents_off + 43 * k + 5, a.parent_cids + 38 * k);
    if (!same || ch.height != a.child_epoch) return;                           // Ok(false) for every proof
    if (a.n_parents == 0) { report_error(a.err, ST_VERIFY, 0, DC_DECODE, CE_RANGE); return; }   // parent_cids[0] panics in the reference
    int32_t pb0 = store_lookup(s, a.parent_cids);
    if (pb0 < 0) { report_error(a.err, ST_VERIFY, 0, DC_MISSING, 2); return; }
    {
        uint32_t pl;
        const uint8_t* pp = store_block(s, (uint32_t)pb0, pl);
        Rd pr(pp, pl);
        HeaderFields ph;
        header_fields(pr, ph);
        if (pr.err) { report_error(a.err, ST_VERIFY, 0, DC_DECODE, pr.err); return; }
        if (ph.height != a.parent_epoch) return;
    }
    *a.consistent = 1;
    // reconstruct_execution_order: every parent header's `messages`; collect_exec_list(verify_txmeta = true)
    for (uint32_t k = 0; k < a.n_parents; k++) {
        int32_t pb = store_lookup(s, a.parent_cids + 38 * k);
        if (pb < 0) { report_error(a.err, ST_VERIFY, 0, DC_MISSING, 3); return; }
        uint32_t pl;
        const uint8_t* pp = store_block(s, (uint32_t)pb, pl);
        Rd pr(pp, pl);
        HeaderFields ph;
        header_fields(pr, ph);
        if (pr.err) { report_error(a.err, ST_VERIFY, 0, DC_DECODE, pr.err); return; }
        for (int q = 0; q < 38; q++) a.txmeta_cids[38 * k + q] = pp[ph.messages_off + q];
    }
    int32_t rb = store_lookup(s, cp + ch.receipts_off);
    if (rb >= 0) *a.receipts_root_blk = (uint32_t)rb;
}
// put_cbor(&(bls_root, secp_root), Blake2b256) == tx_cid (utils.rs:64-73): the strict decoder accepts only the canonical encoding, so
// re-encoding the decoded pair gives the block's own bytes — the recomputed CID is `dag-cbor | blake2b-256 | Blake2b(block)`
__device__ __forceinline__ void verify_txmeta_item(const StoreView& s, const uint8_t* txmeta_cids, uint32_t k, unsigned long long* err) {
    const uint8_t* cid = txmeta_cids + 38 * k;
    int32_t b = store_lookup(s, cid);
    if (b < 0) return;                                                          // reported as missing TxMeta by the walk
    uint32_t len;
    const uint8_t* p = store_block(s, (uint32_t)b, len);
    Rd r(p, len);
    rd_array_exact(r, 2);
    (void)rd_cid(r); (void)rd_cid(r);
    rd_end(r);
    if (r.err) return;                                                          // reported as a decode error by the walk
    static const uint8_t want[6] = {0x01, 0x71, 0xa0, 0xe4, 0x02, 0x20};
    bool ok = true;
    for (int q = 0; q < 6; q++) ok &= cid[q] == want[q];
    Digest d;
    blake2b256(p, len, d);
    Digest c = load_digest(cid + 6);
    if (!ok || !digest_eq(d, c)) report_error(err, ST_VERIFY, 0, DC_CID_MISMATCH, (uint32_t)k);
}

// Amtv0<Receipt>.get(i) WITHOUT recording; returns 1 Some (events root offset in *ev_off, 0xffffffff = None), 0 None, <0 -DevCode
static __device__ int receipts_get_value(const StoreView& s, uint32_t root_blk, uint64_t i, const uint8_t** blk_out, uint32_t* ev_off, uint32_t* detail) {
    uint32_t len;
    const uint8_t* p = store_block(s, root_blk, len);
    Rd r(p, len);
    uint32_t bw, height;
    uint64_t cnt;
    amt_root_begin(r, 0, bw, height, cnt);
    if (r.err) { *detail = r.err; return -(int)DC_DECODE; }
    bool in_range = i < pow_sat(3, height + 1);
    uint32_t lvl = height;
    // (the root node is decoded by `load` whatever the index)
    for (;;) {
        AmtNodeHdr h;
        amt_node_begin(r, 3, h);
        uint32_t nv = rd_array(r);
        uint32_t idx = (uint32_t)((i / pow_sat(3, lvl)) & 7);
        uint32_t want = bm_test(h.bm, idx) ? bm_rank(h.bm, idx) : 0xffffffffu;
        uint32_t found_off = 0xfffffffeu;
        for (uint32_t v = 0; v < nv && !r.err; v++) {
            rd_array_exact(r, 4);
            uint64_t ec = rd_uint(r);
            if (!r.err && ec > 0xffffffffull) rd_fail(r, CE_RANGE);
            uint32_t l;
            (void)rd_bytes(r, l);
            (void)rd_uint(r);
            uint32_t eo = rd_opt_cid(r);
            if (v == want) found_off = eo;
        }
        amt_node_finish(r, h, nv, lvl);
        if (r.err) { *detail = r.err; return -(int)DC_DECODE; }
        if (!in_range) return 0;
        if (h.nl == 0) {
            if (lvl != 0 || want == 0xffffffffu) return 0;
            *blk_out = p; *ev_off = found_off;
            return 1;
        }
        if (want == 0xffffffffu) return 0;
        int32_t child = store_lookup(s, p + h.links_off + 43 * want + 5);
        if (child < 0) { *detail = 0; return -(int)DC_MISSING; }
        p = store_block(s, (uint32_t)child, len);
        r = Rd(p, len);
        lvl--;
    }
}
// Amt<StampedEvent>(v3).get(j): 1 Some (event in ev, its block in *blk_out), 0 None, <0 -DevCode
static __device__ int events_get_value(const StoreView& s, uint32_t root_blk, uint64_t j, const uint8_t** blk_out, EvLog& ev, uint32_t* detail) {
    uint32_t len;
    const uint8_t* p = store_block(s, root_blk, len);
    Rd r(p, len);
    uint32_t bw, height;
    uint64_t cnt;
    amt_root_begin(r, 3, bw, height, cnt);
    if (r.err) { *detail = r.err; return -(int)DC_DECODE; }
    const bool in_range = j < pow_sat(bw, height + 1);
    uint32_t lvl = height;
    for (;;) {
        AmtNodeHdr h;
        amt_node_begin(r, bw, h);
        uint32_t nv = rd_array(r);
        const uint32_t width_mask = (1u << bw) - 1u;
        uint32_t idx = (uint32_t)((j / pow_sat(bw, lvl)) & width_mask);
        uint32_t want = bm_test(h.bm, idx) ? bm_rank(h.bm, idx) : 0xffffffffu;
        bool got = false;
        for (uint32_t v = 0; v < nv && !r.err; v++) {
            EvLog e;
            decode_stamped_event(r, e);
            if (v == want && !r.err) { ev = e; got = true; }
        }
        amt_node_finish(r, h, nv, lvl);
        if (r.err) { *detail = r.err; return -(int)DC_DECODE; }
        if (!in_range) return 0;
        if (h.nl == 0) {
            if (lvl != 0 || !got) return 0;
            *blk_out = p;
            return 1;
        }
        if (want == 0xffffffffu) return 0;
        int32_t child = store_lookup(s, p + h.links_off + 43 * want + 5);
        if (child < 0) { *detail = 0; return -(int)DC_MISSING; }
        p = store_block(s, (uint32_t)child, len);
        r = Rd(p, len);
        lvl--;
    }
}

struct VerifyEventArgs {
    StoreView store;
    const ipcfp_event_proof* proofs;
    uint64_t n;
    const uint8_t* blob;
    uint64_t blob_size;
    const uint32_t* consistent;
    const uint32_t* receipts_root_blk;
    const RawCid* exec_raw;
    const uint32_t* exec_idx;
    uint64_t n_exec;
    const Matcher* filter;        // nullptr: no predicate
    uint8_t* results;
    unsigned long long* err;
};
__device__ __forceinline__ void verify_event_item(const VerifyEventArgs& a, uint64_t t) {
    const StoreView& s = a.store;
    const ipcfp_event_proof& p = a.proofs[t];
    a.results[t] = 0;
    if (!*a.consistent) return;
    // verify_execution_order (:184-204)
    if (p.exec_index >= a.n_exec) return;
    {
        const RawCid c = a.exec_raw[a.exec_idx[p.exec_index]];
        bool eq = true;
        for (int q = 0; q < 6; q++) eq &= p.message_cid[q] == (uint8_t)(c.w[4] >> (8 * q));
        for (int q = 0; q < 32; q++) eq &= p.message_cid[6 + q] == (uint8_t)(c.w[q >> 3] >> (8 * (q & 7)));
        if (!eq) return;
    }
    // verify_receipt_and_event (:207-254)
    if (*a.receipts_root_blk == 0xffffffffu) { report_error(a.err, ST_VERIFY, t, DC_MISSING, 4); return; }
    uint32_t detail = 0, ev_off = 0;
    const uint8_t* rblk = nullptr;
    int got = receipts_get_value(s, *a.receipts_root_blk, p.exec_index, &rblk, &ev_off, &detail);
    if (got < 0) { report_error(a.err, ST_VERIFY, t, (uint32_t)(-got), detail); return; }
    if (got == 0 || ev_off == 0xffffffffu) return;
    int32_t eb = store_lookup(s, rblk + ev_off);
    if (eb < 0) { report_error(a.err, ST_VERIFY, t, DC_MISSING, 5); return; }
    EvLog ev;
    const uint8_t* eblk = nullptr;
    got = events_get_value(s, (uint32_t)eb, p.event_index, &eblk, ev, &detail);
    if (got < 0) { report_error(a.err, ST_VERIFY, t, (uint32_t)(-got), detail); return; }
    if (got == 0) return;
    // verify_event_data_matches (:257-290)
    if (ev.emitter != p.emitter || !ev.some || ev.ntopics != p.n_topics) return;
    if (p.topics_off > a.blob_size || 32ull * p.n_topics > a.blob_size - p.topics_off || p.data_off > a.blob_size || p.data_len > a.blob_size - p.data_off) return;
    for (uint32_t k = 0; k < ev.ntopics; k++) {
        const uint8_t* x = eblk + topic_offset(ev, k);
        const uint8_t* y = a.blob + p.topics_off + 32 * k;
        for (int q = 0; q < 32; q++) if (x[q] != y[q]) return;
    }
    if (ev.data_len != p.data_len) return;
    for (uint32_t q = 0; q < ev.data_len; q++) if (eblk[ev.data_off + q] != a.blob[p.data_off + q]) return;
    if (a.filter) {   // the optional semantic check: matches_log of the spec (events/generator.rs:38-40)
        if (ev.ntopics < 2) return;
        const uint32_t o0 = ev.toff[0], o1 = ev.case_a ? ev.toff[0] + 32 : ev.toff[1];
        if (!(eq32(eblk + o0, a.filter->t0) && eq32(eblk + o1, a.filter->t1))) return;
    }
    a.results[t] = 1;
}

struct VerifyStorageArgs {
    StoreView store;
    const uint8_t* child_cid;
    const uint8_t* state_root_json;   // StorageProof.parent_state_root (the caller's tipset)
    const ipcfp_storage_proof* proofs;
    uint64_t n;
    uint8_t* results;
    unsigned long long* err;
};
__device__ __forceinline__ void verify_storage_item(const VerifyStorageArgs& a, uint64_t t) {
    const StoreView& s = a.store;
    const ipcfp_storage_proof& p = a.proofs[t];
    a.results[t] = 0;
    Recorder rec{nullptr, 0, nullptr, false};   // the verifier records nothing
    // verify_parent_state_root (:98-114)
    int32_t hb = store_lookup(s, a.child_cid);
    if (hb < 0) { report_error(a.err, ST_VERIFY, t, DC_MISSING, 1); return; }
    uint32_t hl;
    const uint8_t* hp = store_block(s, (uint32_t)hb, hl);
    Rd hr(hp, hl);
    uint32_t psr_off = header_parent_state_root(hr);
    if (hr.err) { report_error(a.err, ST_VERIFY, t, DC_DECODE, hr.err); return; }
    const uint8_t* psr = hp + psr_off;
    if (!cid38_equal(psr, a.state_root_json)) return;
    // verify_actor_state (:117-132): get_actor_state (common/decode.rs:17-42)
    int32_t sb = store_lookup(s, psr);
    if (sb < 0) { report_error(a.err, ST_VERIFY, t, DC_MISSING, 2); return; }
    uint32_t sl;
    const uint8_t* sp = store_block(s, (uint32_t)sb, sl);
    Rd sr(sp, sl);
    rd_array_exact(sr, 3);
    uint64_t ver = rd_uint(sr);
    if (!sr.err && ver > 5) rd_fail(sr, CE_RANGE);
    uint32_t actors_off = rd_cid(sr);
    (void)rd_cid(sr);
    rd_end(sr);
    if (sr.err) { report_error(a.err, ST_VERIFY, t, DC_DECODE, sr.err); return; }
    uint8_t key[11];
    uint32_t kl = 0;
    key[kl++] = 0;
    uint64_t id = p.actor_id;
    while (id >= 0x80) { key[kl++] = (uint8_t)(id | 0x80); id >>= 7; }
    key[kl++] = (uint8_t)id;
    bool found;
    ValueRef vr;
    Fail f{0, 0};
    if (!hamt_get(s, rec, sp + actors_off, 5, HV_ACTOR_STATE, key, kl, found, vr, f)) { report_error(a.err, ST_VERIFY, t, f.code, f.detail); return; }
    if (!found) { report_error(a.err, ST_VERIFY, t, DC_ACTOR_NOT_FOUND, 0); return; }
    uint32_t abl;
    const uint8_t* abp = store_block(s, vr.blk, abl);
    Rd ar(abp, abl);
    ar.pos = vr.off;
    uint32_t state_off;
    parse_actor_state(ar, state_off);
    const uint8_t* state_cid = abp + state_off;
    if (!cid38_equal(state_cid, p.actor_state_cid)) return;
    // verify_storage_root (:135-150)
    int32_t eb = store_lookup(s, state_cid);
    if (eb < 0) { report_error(a.err, ST_VERIFY, t, DC_MISSING, 3); return; }
    uint32_t el;
    const uint8_t* ep = store_block(s, (uint32_t)eb, el);
    uint32_t cs_off;
    if (!try_evm_state(ep, el, 6, cs_off) && !try_evm_state(ep, el, 5, cs_off)) { report_error(a.err, ST_VERIFY, t, DC_DECODE, CE_FIELD); return; }
    const uint8_t* storage_root = ep + cs_off;
    if (!cid38_equal(storage_root, p.storage_root)) return;
    // verify_storage_value (:153-170)
    SlotValue sv;
    if (!read_storage_slot(s, rec, storage_root, p.slot, sv, f)) { report_error(a.err, ST_VERIFY, t, f.code, f.detail); return; }
    for (int q = 0; q < 32; q++) if (sv.v32[q] != p.value[q]) return;
    a.results[t] = 1;
}

}  // namespace ipcfp
