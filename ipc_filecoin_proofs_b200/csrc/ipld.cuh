// ipld.cuh — device-side decoders for the IPLD structures on the hot path:
//   AMT nodes  (fvm_ipld_amt 0.7 [UPSTREAM]: node = [bmap, [links], [values]];
//               root v0 = [height,count,node] bw 3; root v3 = [bit_width,height,count,node])
//   StampedEvent / ActorEvent / Entry + extract_evm_log (reference common/evm.rs:13-59)
//   Receipt 4-tuple, HAMT v3 nodes (fvm_ipld_hamt 0.10 [UPSTREAM]) and chain objects.
// One thread decodes one node; the decode contract is DESIGN.md §3.
#pragma once
#include "cbor.cuh"
#include "store.cuh"

namespace ipcfp {

// (actor_id, topic_0, topic_1) — EventMatcher of reference events/generator.rs:23-41
struct Matcher {
    uint64_t t0[4];   // keccak256(event_signature), little-endian word loads of the 32 bytes
    uint64_t t1[4];   // ascii_to_bytes32(topic_1)
    uint64_t actor;
    uint32_t has_actor;
};

__device__ __forceinline__ uint64_t pow_sat(uint32_t bw, uint32_t exp) {
    uint32_t s = bw * exp;
    return s >= 64 ? 0xFFFFFFFFFFFFFFFFull : (1ull << s);
}

__device__ __forceinline__ bool eq32(const uint8_t* p, const uint64_t w[4]) {
    if (load_u64_le(p) != w[0]) return false;
    return load_u64_le(p + 8) == w[1] && load_u64_le(p + 16) == w[2] && load_u64_le(p + 24) == w[3];
}

__device__ __forceinline__ bool cid38_equal(const uint8_t* a, const uint8_t* b) {
    bool eq = true;
    for (int k = 0; k < 38; k++) eq &= a[k] == b[k];
    return eq;
}

// ------------------------------------------------------------------ AMT node framing
struct AmtNodeHdr {
    uint64_t bm[4];      // bitmap bits 0..255 (bit i of the node ↔ bm[i/64] >> (i%64))
    uint32_t pc;         // popcount
    uint32_t nl;         // number of links
    uint32_t links_off;  // offset of the first link item (each exactly 43 bytes)
};
// Reads `[bmap, [links…` up to and including the links array; links are validated.
__device__ __forceinline__ void amt_node_begin(Rd& r, uint32_t bw, AmtNodeHdr& h) {
    rd_array_exact(r, 3);
    uint32_t blen;
    uint32_t boff = rd_bytes(r, blen);
    uint32_t want = bw <= 3 ? 1u : (1u << (bw - 3));
    h.bm[0] = h.bm[1] = h.bm[2] = h.bm[3] = 0;
    h.pc = 0;
    if (!r.err && blen != want) rd_fail(r, CE_AMT);
    if (!r.err) {
        for (uint32_t i = 0; i < blen; i++) h.bm[i >> 3] |= (uint64_t)r.p[boff + i] << (8 * (i & 7));
        uint32_t width = 1u << bw;
        if (width < 8 && (h.bm[0] >> width)) rd_fail(r, CE_AMT);  // bit beyond the node width
        h.pc = (uint32_t)(__popcll(h.bm[0]) + __popcll(h.bm[1]) + __popcll(h.bm[2]) + __popcll(h.bm[3]));
    }
    h.nl = rd_array(r);
    h.links_off = r.pos;
    for (uint32_t k = 0; k < h.nl && !r.err; k++) (void)rd_cid(r);
}
// after the values array has been consumed by the caller
__device__ __forceinline__ void amt_node_finish(Rd& r, const AmtNodeHdr& h, uint32_t nv, uint32_t height) {
    if (r.err) return;
    if (h.nl && nv) { rd_fail(r, CE_AMT); return; }
    if (h.nl) { if (height == 0 || h.pc != h.nl) rd_fail(r, CE_AMT); }
    else { if ((nv && height != 0) || h.pc != nv) rd_fail(r, CE_AMT); }
    if (!r.err) rd_end(r);
}
__device__ __forceinline__ bool bm_test(const uint64_t bm[4], uint32_t i) { return (bm[i >> 6] >> (i & 63)) & 1; }
__device__ __forceinline__ uint32_t bm_rank(const uint64_t bm[4], uint32_t i) {  // set bits below i
    uint32_t c = 0;
    uint32_t w = i >> 6;
    for (uint32_t k = 0; k < w; k++) c += (uint32_t)__popcll(bm[k]);
    uint32_t b = i & 63;
    if (b) c += (uint32_t)__popcll(bm[w] & ((1ull << b) - 1));
    return c;
}
__device__ __forceinline__ uint32_t bm_select(const uint64_t bm[4], uint32_t k) {  // position of the k-th set bit
    for (uint32_t w = 0; w < 4; w++) {
        uint32_t c = (uint32_t)__popcll(bm[w]);
        if (k < c) {
            uint64_t x = bm[w];
            for (uint32_t j = 0; j < k; j++) x &= x - 1;
            return w * 64 + (uint32_t)(__ffsll((long long)x) - 1);
        }
        k -= c;
    }
    return 0xffffffffu;
}
// AMT roots. version 0: [height,count,node] (bw 3); version 3: [bw,height,count,node]
__device__ __forceinline__ void amt_root_begin(Rd& r, int version, uint32_t& bw, uint32_t& height, uint64_t& count) {
    if (version == 0) { rd_array_exact(r, 3); bw = 3; }
    else {
        rd_array_exact(r, 4);
        uint64_t b = rd_uint(r);
        if (!r.err && (b < 1 || b > 8)) rd_fail(r, CE_AMT);
        bw = r.err ? 3 : (uint32_t)b;
    }
    uint64_t h = rd_uint(r);
    if (!r.err && h * bw > 64) rd_fail(r, CE_AMT);
    height = r.err ? 0 : (uint32_t)h;
    count = rd_uint(r);
}

// ------------------------------------------------------------------ StampedEvent + extract_evm_log
struct EvLog {
    uint64_t emitter;
    uint32_t some;       // extract_evm_log returned Some
    uint32_t case_a;     // `topics`/`data` encoding
    uint32_t ntopics;
    uint32_t toff[4];    // Case B: offsets of t1..t4 values; Case A: toff[0] = offset of the topics blob
    uint32_t data_off, data_len;
};
// Decodes one StampedEvent = [emitter, [[flags,key,codec,value]…]] and evaluates
// extract_evm_log (common/evm.rs:13-59) on the fly: last duplicate key wins (:14-17); `topics`
// selects Case A (:20-30); Case B walks t1..t4, any present tK with len != 32 voids the log
// (:45-47), stops at the first gap (:50-52), no t1 ⇒ None (:54-56).
__device__ __forceinline__ void parse_stamped_event(Rd& r, EvLog& ev) {
    rd_array_exact(r, 2);
    ev.emitter = rd_uint(r);
    uint32_t ne = rd_array(r);
    uint32_t have = 0;  // bit0..3: t1..t4, bit4: d, bit5: topics, bit6: data
    uint32_t t_off0 = 0, t_off1 = 0, t_off2 = 0, t_off3 = 0, len_ok = 0;
    uint32_t d_off = 0, d_len = 0, tp_off = 0, tp_len = 0, da_off = 0, da_len = 0;
    for (uint32_t e = 0; e < ne && !r.err; e++) {
        rd_array_exact(r, 4);
        (void)rd_uint(r);
        uint32_t klen, vlen;
        uint32_t koff = rd_text(r, klen);
        (void)rd_uint(r);
        uint32_t voff = rd_bytes(r, vlen);
        if (r.err) break;
        const uint8_t* k = r.p + koff;
        if (klen == 2 && k[0] == 't') {
            uint32_t idx = (uint32_t)k[1] - (uint32_t)'1';
            if (idx < 4) {
                have |= 1u << idx;
                if (vlen == 32) len_ok |= 1u << idx; else len_ok &= ~(1u << idx);
                if (idx == 0) t_off0 = voff; else if (idx == 1) t_off1 = voff; else if (idx == 2) t_off2 = voff; else t_off3 = voff;
            }
        } else if (klen == 1 && k[0] == 'd') { have |= 16; d_off = voff; d_len = vlen; }
        else if (klen == 6 && bytes_eq(k, "topics", 6)) { have |= 32; tp_off = voff; tp_len = vlen; }
        else if (klen == 4 && bytes_eq(k, "data", 4)) { have |= 64; da_off = voff; da_len = vlen; }
    }
    ev.some = 0; ev.case_a = 0; ev.ntopics = 0; ev.data_off = 0; ev.data_len = 0;
    ev.toff[0] = ev.toff[1] = ev.toff[2] = ev.toff[3] = 0;
    if (r.err) return;
    if (have & 32) {
        ev.case_a = 1;
        if (tp_len % 32 == 0) {
            ev.some = 1; ev.ntopics = tp_len / 32; ev.toff[0] = tp_off;
            if (have & 64) { ev.data_off = da_off; ev.data_len = da_len; }
        }
        return;
    }
    uint32_t n = 0;
    bool none = false;
#pragma unroll
    for (uint32_t i = 0; i < 4; i++) {
        if (none || n != i) break;          // stopped at an earlier gap
        if (!(have & (1u << i))) break;
        if (!(len_ok & (1u << i))) { none = true; break; }
        n++;
    }
    if (none || n == 0) return;
    ev.some = 1; ev.ntopics = n;
    ev.toff[0] = t_off0; ev.toff[1] = t_off1; ev.toff[2] = t_off2; ev.toff[3] = t_off3;
    if (have & 16) { ev.data_off = d_off; ev.data_len = d_len; }
}
// actor filter (events/generator.rs:220-224) then matches_log (:38-40)
__device__ __forceinline__ bool event_matches(const uint8_t* p, const EvLog& ev, const Matcher& m) {
    if (m.has_actor && ev.emitter != m.actor) return false;
    if (!ev.some || ev.ntopics < 2) return false;
    uint32_t o0 = ev.toff[0], o1 = ev.case_a ? ev.toff[0] + 32 : ev.toff[1];
    return eq32(p + o0, m.t0) && eq32(p + o1, m.t1);
}
__device__ __forceinline__ uint32_t topic_offset(const EvLog& ev, uint32_t k) { return ev.case_a ? ev.toff[0] + 32 * k : ev.toff[k]; }

// ------------------------------------------------------------------ Receipt = [exit_code, return_data, gas_used, events_root|null]
__device__ __forceinline__ void parse_receipt(Rd& r) {
    rd_array_exact(r, 4);
    uint64_t ec = rd_uint(r);
    if (!r.err && ec > 0xffffffffull) rd_fail(r, CE_RANGE);
    uint32_t l;
    (void)rd_bytes(r, l);
    (void)rd_uint(r);
    (void)rd_opt_cid(r);
}

// ------------------------------------------------------------------ HAMT (fvm_ipld_hamt v3 layout)
enum HamtValueKind { HV_ACTOR_STATE = 0, HV_U8VEC = 1 };
// value decoders: validate and remember where the value starts
__device__ __forceinline__ void parse_actor_state(Rd& r, uint32_t& state_cid_off) {
    rd_array_exact(r, 5);
    (void)rd_cid(r);
    state_cid_off = rd_cid(r);
    (void)rd_uint(r);
    uint32_t l;
    (void)rd_bytes(r, l);
    if (rd_peek_null(r)) r.pos++; else (void)rd_bytes(r, l);
}
// serde Vec<u8>: CBOR array of u8; returns element count, elements start at r.pos after the head
__device__ __forceinline__ uint32_t parse_u8vec(Rd& r, uint32_t& first_elem_off) {
    uint32_t n = rd_array(r);
    first_elem_off = r.pos;
    for (uint32_t i = 0; i < n && !r.err; i++) { uint64_t x = rd_uint(r); if (!r.err && x > 255) rd_fail(r, CE_RANGE); }
    return n;
}

struct HamtHit {
    int32_t kind;          // 0 = None, 1 = value found, 2 = follow link
    uint32_t val_off;      // offset of the value item (kind 1)
    uint32_t link_off;     // offset of the 38 CID bytes (kind 2)
};
// Decodes a whole HAMT node (all pointers, all buckets, every value — like serde does) and resolves
// slot `idx` for `key`.
__device__ __forceinline__ void hamt_node_lookup(Rd& r, int vkind, uint32_t idx, const uint8_t* key, uint32_t keylen, HamtHit& hit) {
    hit.kind = 0; hit.val_off = 0; hit.link_off = 0;
    rd_array_exact(r, 2);
    uint32_t blen;
    uint32_t boff = rd_bytes(r, blen);
    if (!r.err && blen > 32) rd_fail(r, CE_HAMT);
    uint64_t bf[4] = {0, 0, 0, 0};  // bf bit i ↔ child i; big-endian byte string, right aligned
    if (!r.err) for (uint32_t i = 0; i < blen; i++) { uint32_t bytepos = blen - 1 - i; bf[i >> 3] |= (uint64_t)r.p[boff + bytepos] << (8 * (i & 7)); }
    uint32_t np = rd_array(r);
    uint32_t pc = (uint32_t)(__popcll(bf[0]) + __popcll(bf[1]) + __popcll(bf[2]) + __popcll(bf[3]));
    bool present = bm_test(bf, idx);
    uint32_t want = present ? bm_rank(bf, idx) : 0xffffffffu;
    for (uint32_t k = 0; k < np && !r.err; k++) {
        uint32_t mj = rd_peek_major(r);
        if (mj == 6) {
            uint32_t off = rd_cid(r);
            if (k == want && !r.err) { hit.kind = 2; hit.link_off = off; }
        } else if (mj == 4) {
            uint32_t nk = rd_array(r);
            for (uint32_t j = 0; j < nk && !r.err; j++) {
                rd_array_exact(r, 2);
                uint32_t kl;
                uint32_t ko = rd_bytes(r, kl);
                uint32_t voff = r.pos;
                if (vkind == HV_ACTOR_STATE) { uint32_t s; parse_actor_state(r, s); }
                else { uint32_t f; (void)parse_u8vec(r, f); }
                if (k == want && !r.err && hit.kind == 0 && kl == keylen) {
                    bool eq = true;
                    for (uint32_t b = 0; b < kl; b++) eq &= r.p[ko + b] == key[b];
                    if (eq) { hit.kind = 1; hit.val_off = voff; }
                }
            }
        } else if (!r.err) rd_fail(r, CE_HAMT);
    }
    if (!r.err) rd_end(r);
    if (!r.err && pc != np) rd_fail(r, CE_HAMT);
    if (r.err) hit.kind = 0;
}
// bits [consumed, consumed+bw) of a SHA-256 digest given as 8 big-endian words, MSB first
__device__ __forceinline__ uint32_t hash_bits(const uint32_t h_be[8], uint32_t consumed, uint32_t bw) {
    uint32_t v = 0;
    for (uint32_t k = 0; k < bw; k++) {
        uint32_t bit = consumed + k;
        v = (v << 1) | ((h_be[bit >> 5] >> (31 - (bit & 31))) & 1u);
    }
    return v;
}

}  // namespace ipcfp
