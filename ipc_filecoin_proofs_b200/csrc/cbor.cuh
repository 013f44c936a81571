// cbor.cuh — streaming, allocation-free strict DAG-CBOR reader for device code.
//
// Restates the decode behaviour the reference gets from serde_ipld_dagcbor 0.6 / fvm_ipld_encoding
// ([UPSTREAM], DESIGN.md §3 "decode contract"): definite lengths only, minimal-length heads, typed
// positions, exact tuple lengths, tag 42 only, no trailing bytes. Errors are sticky: the first
// failure is latched in Rd::err, the cursor jumps to the end and later reads return zeros, so
// callers may check once per node instead of after every item (keeps warps converged).
#pragma once
#include "common.cuh"

namespace ipcfp {

enum CborErr : uint32_t {
    CE_EOF = 1, CE_AI = 2, CE_NONMIN = 3, CE_TYPE = 4, CE_BOUNDS = 5, CE_UTF8 = 6, CE_LEN = 7, CE_CID = 8, CE_TRAIL = 9,
    CE_AMT = 10, CE_HAMT = 11, CE_TAG = 12, CE_SIMPLE = 13, CE_RANGE = 14, CE_FIELD = 15
};

struct Rd {
    const uint8_t* p;
    uint32_t n;
    uint32_t pos;
    uint32_t err;
    __device__ __forceinline__ Rd(const uint8_t* ptr, uint32_t len) : p(ptr), n(len), pos(0), err(0) {}
};

__device__ __forceinline__ void rd_fail(Rd& r, uint32_t code) {
    if (!r.err) r.err = code;
    r.pos = r.n;
}

// Decodes one head. Returns major (0..7); arg and additional-info in out params.
__device__ __forceinline__ uint32_t rd_head(Rd& r, uint64_t& arg, uint32_t& ai) {
    arg = 0; ai = 0;
    if (r.pos >= r.n) { rd_fail(r, CE_EOF); return 0xff; }
    uint32_t ib = r.p[r.pos];
    uint32_t major = ib >> 5;
    ai = ib & 31;
    if (ai < 24) { arg = ai; r.pos += 1; return major; }
    if (ai > 27) { rd_fail(r, CE_AI); return 0xff; }
    uint32_t nb = 1u << (ai - 24);
    if (r.n - r.pos - 1 < nb) { rd_fail(r, CE_EOF); return 0xff; }
    uint64_t v = 0;
    for (uint32_t i = 0; i < nb; i++) v = (v << 8) | r.p[r.pos + 1 + i];
    arg = v;
    r.pos += 1 + nb;
    if (major != 7) {
        uint64_t minv = ai == 24 ? 24ull : (ai == 25 ? 0x100ull : (ai == 26 ? 0x10000ull : 0x100000000ull));
        if (v < minv) { rd_fail(r, CE_NONMIN); return 0xff; }
    }
    return major;
}

__device__ __forceinline__ uint64_t rd_uint(Rd& r) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    if (m != 0) { rd_fail(r, CE_TYPE); return 0; }
    return a;
}
__device__ __forceinline__ int64_t rd_int(Rd& r) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    if (m > 1 || a > 0x7fffffffffffffffull) { rd_fail(r, CE_TYPE); return 0; }
    return m == 0 ? (int64_t)a : -1 - (int64_t)a;
}
// byte string: returns offset of payload, len in out param
__device__ __forceinline__ uint32_t rd_bytes(Rd& r, uint32_t& len) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    len = 0;
    if (m != 2) { rd_fail(r, CE_TYPE); return r.n; }
    if (a > (uint64_t)(r.n - r.pos)) { rd_fail(r, CE_BOUNDS); return r.n; }
    uint32_t off = r.pos;
    len = (uint32_t)a;
    r.pos += len;
    return off;
}
// Rust str::from_utf8 rules
static __device__ __noinline__ bool utf8_valid(const uint8_t* s, uint32_t len) {
    uint32_t i = 0;
    while (i < len) {
        uint32_t c = s[i];
        if (c < 0x80) { i++; continue; }
        uint32_t extra, cp;
        if ((c & 0xe0) == 0xc0) { extra = 1; cp = c & 0x1f; }
        else if ((c & 0xf0) == 0xe0) { extra = 2; cp = c & 0x0f; }
        else if ((c & 0xf8) == 0xf0) { extra = 3; cp = c & 0x07; }
        else return false;
        if (extra > len - 1 - i) return false;
        for (uint32_t k = 1; k <= extra; k++) { uint32_t d = s[i + k]; if ((d & 0xc0) != 0x80) return false; cp = (cp << 6) | (d & 0x3f); }
        if (extra == 1 && cp < 0x80) return false;
        if (extra == 2 && (cp < 0x800 || (cp >= 0xd800 && cp <= 0xdfff))) return false;
        if (extra == 3 && (cp < 0x10000 || cp > 0x10ffff)) return false;
        i += 1 + extra;
    }
    return true;
}
// text string: returns offset, len; validates UTF-8
__device__ __forceinline__ uint32_t rd_text(Rd& r, uint32_t& len) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    len = 0;
    if (m != 3) { rd_fail(r, CE_TYPE); return r.n; }
    if (a > (uint64_t)(r.n - r.pos)) { rd_fail(r, CE_BOUNDS); return r.n; }
    uint32_t off = r.pos;
    len = (uint32_t)a;
    bool ascii = true;
    for (uint32_t i = 0; i < len; i++) ascii &= r.p[off + i] < 0x80;
    if (!ascii && !utf8_valid(r.p + off, len)) { rd_fail(r, CE_UTF8); return r.n; }
    r.pos += len;
    return off;
}
// array head; the count is bounded by the bytes left (every item takes ≥ 1 byte)
__device__ __forceinline__ uint32_t rd_array(Rd& r) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    if (m != 4) { rd_fail(r, CE_TYPE); return 0; }
    if (a > (uint64_t)(r.n - r.pos)) { rd_fail(r, CE_LEN); return 0; }
    return (uint32_t)a;
}
__device__ __forceinline__ void rd_array_exact(Rd& r, uint32_t k) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    if (m != 4 || a != k) rd_fail(r, m != 4 ? CE_TYPE : CE_LEN);
}
__device__ __forceinline__ uint32_t rd_map(Rd& r) {
    uint64_t a; uint32_t ai;
    uint32_t m = rd_head(r, a, ai);
    if (m != 5) { rd_fail(r, CE_TYPE); return 0; }
    if (a > (uint64_t)(r.n - r.pos)) { rd_fail(r, CE_LEN); return 0; }
    return (uint32_t)a;
}
__device__ __forceinline__ bool rd_peek_null(const Rd& r) { return r.pos < r.n && r.p[r.pos] == 0xf6; }
__device__ __forceinline__ uint32_t rd_peek_major(Rd& r) {
    if (r.pos >= r.n) { rd_fail(r, CE_EOF); return 0xff; }
    return r.p[r.pos] >> 5;
}
// tag 42 link: `d8 2a 58 27 00 <38-byte CIDv1>`; returns the offset of the 38 CID bytes
__device__ __forceinline__ uint32_t rd_cid(Rd& r) {
    if (r.n - r.pos < 43) { rd_fail(r, r.pos >= r.n ? CE_EOF : CE_CID); return r.n; }
    const uint8_t* q = r.p + r.pos;
    if (q[0] != 0xd8 || q[1] != 0x2a) {
        // distinguish "not a tag 42" from a malformed head; either way a decode error
        rd_fail(r, CE_TAG);
        return r.n;
    }
    if (q[2] != 0x58 || q[3] != 0x27 || q[4] != 0x00 || q[5] != 0x01) { rd_fail(r, CE_CID); return r.n; }
    uint32_t off = r.pos + 5;
    r.pos += 43;
    return off;
}
// Option<Cid>: returns r.n-sentinel (0xffffffff) when null
__device__ __forceinline__ uint32_t rd_opt_cid(Rd& r) {
    if (rd_peek_null(r)) { r.pos++; return 0xffffffffu; }
    return rd_cid(r);
}
// serde IgnoredAny: any well-formed DAG-CBOR item
static __device__ __noinline__ void rd_skip_any(Rd& r) {
    uint64_t remaining = 1;
    while (remaining && !r.err) {
        remaining--;
        uint64_t a; uint32_t ai;
        uint32_t m = rd_head(r, a, ai);
        switch (m) {
            case 0: case 1: break;
            case 2:
                if (a > (uint64_t)(r.n - r.pos)) rd_fail(r, CE_BOUNDS); else r.pos += (uint32_t)a;
                break;
            case 3:
                if (a > (uint64_t)(r.n - r.pos)) rd_fail(r, CE_BOUNDS);
                else if (!utf8_valid(r.p + r.pos, (uint32_t)a)) rd_fail(r, CE_UTF8);
                else r.pos += (uint32_t)a;
                break;
            case 4: if (a > (uint64_t)(r.n - r.pos)) rd_fail(r, CE_LEN); else remaining += a; break;
            case 5: if (a > (uint64_t)(r.n - r.pos) / 2 + 1) rd_fail(r, CE_LEN); else remaining += 2 * a; break;
            case 6: {
                if (a != 42) { rd_fail(r, CE_TAG); break; }
                uint64_t b; uint32_t bi;
                uint32_t mb = rd_head(r, b, bi);
                if (mb != 2 || b > (uint64_t)(r.n - r.pos) || b < 1 || r.p[r.pos] != 0) rd_fail(r, CE_CID);
                else r.pos += (uint32_t)b;
                break;
            }
            case 7:
                if (ai == 20 || ai == 21 || ai == 22 || ai == 27) break;
                rd_fail(r, CE_SIMPLE);
                break;
            default: break;  // error already latched
        }
    }
}
__device__ __forceinline__ void rd_end(Rd& r) {
    if (!r.err && r.pos != r.n) rd_fail(r, CE_TRAIL);
}

// compare n bytes at p with a constant text (n small)
__device__ __forceinline__ bool bytes_eq(const uint8_t* p, const char* s, uint32_t n) {
    bool eq = true;
    for (uint32_t i = 0; i < n; i++) eq &= p[i] == (uint8_t)s[i];
    return eq;
}

}  // namespace ipcfp
