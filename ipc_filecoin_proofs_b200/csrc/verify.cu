// verify.cu — GPU-batched verifiers over a WITNESS store.
//
//   verify_event_proofs     reference src/proofs/events/verifier.rs:51-290  (verify_event_proof → verify_single_proof per proof)
//   verify_storage_proofs   reference src/proofs/storage/verifier.rs:24-170 (verify_storage_proof per proof)
//
// The witness blocks go into an ipcfp_store first (ipcfp_store_create with IPCFP_STORE_VERIFY_CIDS): that is the Blake2b-256 check of
// EVERY witness block which the reference's load_witness_store leaves out (`put_keyed`, events/verifier.rs:79-89 — SURVEY F6).
// What the reference repeats per proof but depends on the tipset only is done once per call: header consistency (:147-181), the
// TxMeta recompute and the execution order (events/utils.rs:16-30, :64-73 — O(messages) PER PROOF in the reference). Then one warp
// per proof (lane 0 walks, see storage.cu for why) replays verify_execution_order (:184-204: exec[exec_index] == message_cid, which
// for a duplicate-free list is `position(message_cid) == exec_index`), Amtv0<Receipt>.get(exec_index) → events_root →
// Amt<StampedEvent>.get(event_index) (:207-254) and verify_event_data_matches (:257-290).
// Results are the reference's Vec<bool>; an Err of the reference (missing block, decode failure, TxMeta mismatch) fails the call
// with the index of the FIRST proof that meets it. Trust anchors (:124-144) are host-side policy closures and stay with the caller.
#include <algorithm>
#include <cstring>

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
__global__ void k_verify_tipset(VerifyTipsetArgs a) {
    if (threadIdx.x || blockIdx.x) return;
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
__global__ void k_verify_txmeta(StoreView s, const uint8_t* txmeta_cids, uint32_t n_parents, unsigned long long* err) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_parents) return;
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
__global__ void __launch_bounds__(128) k_verify_events(VerifyEventArgs a) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= a.n || (threadIdx.x & 31)) return;
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

static void throw_verify_error(uint64_t key) {
    const uint32_t code = (uint32_t)(key >> 8) & 0xff, detail = (uint32_t)key & 0xff;
    const uint64_t index = (key >> 16) & 0xFFFFFFFFFFull;
    switch (code) {
        case DC_MISSING: throw Error(IPCFP_ERR_MISSING_BLOCK, "missing block in the witness (detail " + std::to_string(detail) + ")", index);
        case DC_CID_MISMATCH: throw Error(IPCFP_ERR_CID_MISMATCH, "TxMeta mismatch: header vs recomputed (parent " + std::to_string(detail) + ")", index);
        case DC_ACTOR_NOT_FOUND: throw Error(IPCFP_ERR_ACTOR_NOT_FOUND, "actor not found", index);
        default: throw Error(IPCFP_ERR_DECODE, "decode error in the witness (detail " + std::to_string(detail) + ")", index);
    }
}

void verify_event_proofs(Store* s, const ipcfp_tipset_desc* t, const ipcfp_event_proof* proofs, uint64_t n, const uint8_t* data_blob, uint64_t blob_size,
                         const ipcfp_event_spec* filter, uint8_t* results) {
    s->use();
    if (!t || !t->child_cid || (t->n_parents && !t->parent_cids)) throw Error(IPCFP_ERR_INVALID_ARG, "tipset descriptor has null fields");
    if (t->n_parents > 64) throw Error(IPCFP_ERR_UNSUPPORTED, "too many parent blocks");
    if (n && (!proofs || !results)) throw Error(IPCFP_ERR_INVALID_ARG, "null argument");
    if (n == 0) return;
    cudaStream_t st = s->stream;
    unsigned long long* dw = s->dev_words.p;
    uint64_t* hw = s->host_words.p;
    const uint32_t P = t->n_parents;
    IPCFP_CUDA(cudaMemsetAsync(dw, 0xff, 8, st));
    AsyncBuf<uint8_t> d_cids(38ull * (2 * P + 2) + 64, st), d_blob(blob_size + 16, st), d_res(n + 16, st);
    AsyncBuf<uint32_t> d_flags(8, st);
    AsyncBuf<ipcfp_event_proof> d_proofs(n, st);
    IPCFP_CUDA(cudaMemcpyAsync(d_cids.p, t->child_cid, 38, cudaMemcpyHostToDevice, st));
    if (P) IPCFP_CUDA(cudaMemcpyAsync(d_cids.p + 38, t->parent_cids, 38ull * P, cudaMemcpyHostToDevice, st));
    IPCFP_CUDA(cudaMemcpyAsync(d_proofs.p, proofs, n * sizeof(ipcfp_event_proof), cudaMemcpyHostToDevice, st));
    if (blob_size) IPCFP_CUDA(cudaMemcpyAsync(d_blob.p, data_blob, blob_size, cudaMemcpyHostToDevice, st));
    uint8_t* d_tx = d_cids.p + 38ull * (P + 1);
    VerifyTipsetArgs ta;
    ta.store = s->view; ta.parent_cids = d_cids.p + 38; ta.child_cid = d_cids.p; ta.n_parents = P;
    ta.parent_epoch = t->parent_epoch; ta.child_epoch = t->child_epoch;
    ta.consistent = d_flags.p; ta.receipts_root_blk = d_flags.p + 1; ta.txmeta_cids = d_tx; ta.err = dw;
    k_verify_tipset<<<1, 32, 0, st>>>(ta); IPCFP_LAUNCH_CHECK();
    std::vector<uint8_t> h_tx(38ull * P + 8);
    uint32_t h_flags[2] = {0, 0};
    if (P) IPCFP_CUDA(cudaMemcpyAsync(h_tx.data(), d_tx, 38ull * P, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaMemcpyAsync(h_flags, d_flags.p, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaMemcpyAsync(hw, dw, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    if (hw[0] != IPCFP_NO_ERROR) throw_verify_error(hw[0]);
    ExecOrderOut exo;
    if (h_flags[0]) {
        // collect_exec_list(verify_txmeta = true) once for the whole batch: TxMeta recompute, then the engine's own message-AMT walk
        // + first-seen dedup (the same kernels generate_event_proof uses), on the TxMeta links taken from the parent HEADERS
        k_verify_txmeta<<<div_up(P, 64), 64, 0, st>>>(s->view, d_tx, P, dw); IPCFP_LAUNCH_CHECK();
        TipsetDev td;
        td.parent_epoch = t->parent_epoch; td.child_epoch = t->child_epoch; td.n_parents = P;
        td.parent_cids.assign(t->parent_cids, t->parent_cids + 38ull * P);
        td.txmeta_cids.assign(h_tx.begin(), h_tx.begin() + 38ull * P);
        memcpy(td.child_cid, t->child_cid, 38);
        memcpy(td.receipts_root, t->child_cid, 38);   // unused in execution-order-only mode (any CID of the store)
        td.n_receipts = 0;
        td.events_roots.alloc(64);
        td.has_root.alloc(64);
        ipcfp_event_spec dummy;
        memset(&dummy, 0, sizeof dummy);
        dummy.event_signature = "";
        dummy.topic_1 = "";
        IPCFP_CUDA(cudaMemcpyAsync(hw + 40, dw, 8, cudaMemcpyDeviceToHost, st));
        IPCFP_CUDA(cudaStreamSynchronize(st));
        const uint64_t tx_key = hw[40];
        try {
            (void)generate_event_proof(s, nullptr, td, &dummy, IPCFP_SCAN_SKIP_TX_AMTS, false, 0, 0, 1, 0, nullptr, &exo);
        } catch (Error& e) {
            e.index = 0;   // failures of the walk surface at the first proof, like everything that depends on the tipset only
            throw;
        }
        if (tx_key != IPCFP_NO_ERROR) throw_verify_error(tx_key);
        IPCFP_CUDA(cudaMemsetAsync(dw, 0xff, 8, st));
    }
    AsyncBuf<Matcher> d_filter;
    if (filter) {
        if (!filter->event_signature || !filter->topic_1) throw Error(IPCFP_ERR_INVALID_ARG, "filter spec has null fields");
        Matcher m;
        memset(&m, 0, sizeof m);
        // keccak256(signature) on the device through the batched-hash entry (K2)
        uint8_t t0[32];
        uint64_t off0 = 0;
        uint32_t len0 = (uint32_t)strlen(filter->event_signature);
        hash_batch(1, (const uint8_t*)filter->event_signature, len0, &off0, &len0, 1, s->device, t0);
        memcpy(m.t0, t0, 32);
        uint8_t t1[32];
        memset(t1, 0, 32);
        size_t n1 = strlen(filter->topic_1);
        memcpy(t1, filter->topic_1, n1 < 32 ? n1 : 32);
        memcpy(m.t1, t1, 32);
        d_filter.alloc(1, st);
        IPCFP_CUDA(cudaMemcpyAsync(d_filter.p, &m, sizeof m, cudaMemcpyHostToDevice, st));
        IPCFP_CUDA(cudaStreamSynchronize(st));   // m is a stack object
    }
    VerifyEventArgs va;
    va.store = s->view; va.proofs = d_proofs.p; va.n = n; va.blob = d_blob.p; va.blob_size = blob_size;
    va.consistent = d_flags.p; va.receipts_root_blk = d_flags.p + 1;
    va.exec_raw = exo.exec_raw.p; va.exec_idx = exo.exec_idx.p; va.n_exec = exo.n_exec;
    va.filter = filter ? d_filter.p : nullptr; va.results = d_res.p; va.err = dw;
    k_verify_events<<<div_up(n * 32, 128), 128, 0, st>>>(va); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaMemcpyAsync(results, d_res.p, n, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaMemcpyAsync(hw, dw, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    if (hw[0] != IPCFP_NO_ERROR) throw_verify_error(hw[0]);
}

// ------------------------------------------------------------------------------------------ storage
struct VerifyStorageArgs {
    StoreView store;
    const uint8_t* child_cid;
    const uint8_t* state_root_json;   // StorageProof.parent_state_root (the caller's tipset)
    const ipcfp_storage_proof* proofs;
    uint64_t n;
    uint8_t* results;
    unsigned long long* err;
};
__global__ void __launch_bounds__(128) k_verify_storage(VerifyStorageArgs a) {
    const uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= a.n || (threadIdx.x & 31)) return;
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

void verify_storage_proofs(Store* s, const ipcfp_tipset_desc* t, const ipcfp_storage_proof* proofs, uint64_t n, uint8_t* results) {
    s->use();
    if (!t || !t->child_cid || !t->child_parent_state_root) throw Error(IPCFP_ERR_INVALID_ARG, "tipset descriptor lacks child_cid / parent_state_root");
    if (n && (!proofs || !results)) throw Error(IPCFP_ERR_INVALID_ARG, "null argument");
    if (n == 0) return;
    cudaStream_t st = s->stream;
    unsigned long long* dw = s->dev_words.p;
    uint64_t* hw = s->host_words.p;
    IPCFP_CUDA(cudaMemsetAsync(dw, 0xff, 8, st));
    AsyncBuf<uint8_t> d_in(128, st), d_res(n + 16, st);
    AsyncBuf<ipcfp_storage_proof> d_proofs(n, st);
    IPCFP_CUDA(cudaMemcpyAsync(d_in.p, t->child_cid, 38, cudaMemcpyHostToDevice, st));
    IPCFP_CUDA(cudaMemcpyAsync(d_in.p + 64, t->child_parent_state_root, 38, cudaMemcpyHostToDevice, st));
    IPCFP_CUDA(cudaMemcpyAsync(d_proofs.p, proofs, n * sizeof(ipcfp_storage_proof), cudaMemcpyHostToDevice, st));
    VerifyStorageArgs a;
    a.store = s->view; a.child_cid = d_in.p; a.state_root_json = d_in.p + 64; a.proofs = d_proofs.p; a.n = n; a.results = d_res.p; a.err = dw;
    k_verify_storage<<<div_up(n * 32, 128), 128, 0, st>>>(a); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaMemcpyAsync(results, d_res.p, n, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaMemcpyAsync(hw, dw, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    if (hw[0] != IPCFP_NO_ERROR) throw_verify_error(hw[0]);
}

}  // namespace ipcfp
