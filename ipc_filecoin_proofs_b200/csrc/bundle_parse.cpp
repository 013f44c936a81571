// bundle_parse.cpp — the way back of bundle_json.cpp, in the boundary language (plain C++, built with g++, no CUDA).
//
// ipcfp_bundle_from_json reads what `serde_json::from_str::<UnifiedProofBundle>` / `::<EventProofBundle>` reads in the reference
// (src/proofs/common/bundle.rs:10-45, src/proofs/events/bundle.rs:5-30, src/proofs/storage/bundle.rs:5-14) into the PODs the
// batched verifiers take (ipcfp_verify_event_proofs / ipcfp_verify_storage_proofs) and the flat block arrays
// ipcfp_store_create takes for the witness store — so a host that holds a bundle as JSON can verify it through the C ABI alone:
//
//     ipcfp_bundle_from_json(text, len, &pb);
//     ipcfp_store_create(pb->witness.cids, pb->witness.offsets, pb->witness.lengths, pb->witness.blob, pb->witness.blob_size,
//                        pb->witness.n_blocks, device, IPCFP_STORE_VERIFY_CIDS, &ws);
//     ipcfp_verify_event_proofs(ws, &pb->tipset, pb->event_proofs, pb->n_event_proofs, pb->data_blob, pb->data_blob_size, filter, results);
//
// Accepted spellings mirror ipc_filecoin_proofs_b200/bundle_json.py: CID strings are multibase base32 ("b…"); `ProofBlock.cid` may be
// the byte array cid 0.11's Serialize emits, a {"/": "b…"} link or a plain string; hex fields carry "0x"; block data is standard
// base64 with padding. serde ignores unknown fields: so does this parser. The fields every proof of a bundle shares (epochs, parent
// tipset CIDs, child block CID, parent state root) are returned once, as an ipcfp_tipset_desc; a bundle whose proofs disagree on them
// is refused (IPCFP_ERR_UNSUPPORTED) — the C ABI's verifiers take one tipset pair per call, as the generators produce them.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ipcfp.h"

namespace {

struct JV {
    enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
    bool b = false;
    std::string s;   // NUM: the literal's text, STR: the decoded string
    std::vector<JV> a;
    std::vector<std::pair<std::string, JV>> o;
    const JV* get(const char* k) const {
        if (t != OBJ) return nullptr;
        for (auto& kv : o) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct Parser {
    const char* p;
    const char* e;
    bool ok = true;
    void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) p++; }
    bool lit(const char* w) {
        size_t n = strlen(w);
        if ((size_t)(e - p) < n || memcmp(p, w, n)) return false;
        p += n;
        return true;
    }
    static void utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o.push_back((char)c);
        else if (c < 0x800) { o.push_back((char)(0xc0 | (c >> 6))); o.push_back((char)(0x80 | (c & 63))); }
        else if (c < 0x10000) { o.push_back((char)(0xe0 | (c >> 12))); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
        else { o.push_back((char)(0xf0 | (c >> 18))); o.push_back((char)(0x80 | ((c >> 12) & 63))); o.push_back((char)(0x80 | ((c >> 6) & 63))); o.push_back((char)(0x80 | (c & 63))); }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4) return false;
        v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            uint32_t d = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99;
            if (d == 99) return false;
            v = v * 16 + d;
        }
        return true;
    }
    bool str(std::string& out) {
        if (p >= e || *p != '"') return false;
        p++;
        const char* run = p;
        for (;;) {
            if (p >= e) return false;
            unsigned char c = (unsigned char)*p;
            if (c == '"') { out.append(run, p - run); p++; return true; }
            if (c < 0x20) return false;
            if (c != '\\') { p++; continue; }
            out.append(run, p - run);
            p++;
            if (p >= e) return false;
            char x = *p++;
            switch (x) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t v;
                    if (!hex4(v)) return false;
                    if (v >= 0xd800 && v < 0xdc00) {   // surrogate pair
                        uint32_t w;
                        if (e - p < 6 || p[0] != '\\' || p[1] != 'u') return false;
                        p += 2;
                        if (!hex4(w) || w < 0xdc00 || w > 0xdfff) return false;
                        v = 0x10000 + ((v - 0xd800) << 10) + (w - 0xdc00);
                    } else if (v >= 0xdc00 && v <= 0xdfff) return false;
                    utf8(out, v);
                    break;
                }
                default: return false;
            }
            run = p;
        }
    }
    bool value(JV& v, int depth) {
        if (depth > 64) return false;
        ws();
        if (p >= e) return false;
        char c = *p;
        if (c == '{') {
            p++;
            v.t = JV::OBJ;
            ws();
            if (p < e && *p == '}') { p++; return true; }
            for (;;) {
                ws();
                std::string k;
                if (!str(k)) return false;
                ws();
                if (p >= e || *p != ':') return false;
                p++;
                v.o.emplace_back(std::move(k), JV());
                if (!value(v.o.back().second, depth + 1)) return false;
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == '}') { p++; return true; }
                return false;
            }
        }
        if (c == '[') {
            p++;
            v.t = JV::ARR;
            ws();
            if (p < e && *p == ']') { p++; return true; }
            for (;;) {
                v.a.emplace_back();
                if (!value(v.a.back(), depth + 1)) return false;
                ws();
                if (p < e && *p == ',') { p++; continue; }
                if (p < e && *p == ']') { p++; return true; }
                return false;
            }
        }
        if (c == '"') { v.t = JV::STR; return str(v.s); }
        if (c == 't') { v.t = JV::BOOL; v.b = true; return lit("true"); }
        if (c == 'f') { v.t = JV::BOOL; v.b = false; return lit("false"); }
        if (c == 'n') { v.t = JV::NUL; return lit("null"); }
        if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s0 = p;
            if (*p == '-') p++;
            if (p >= e || *p < '0' || *p > '9') return false;
            if (*p == '0') p++; else while (p < e && *p >= '0' && *p <= '9') p++;
            if (p < e && *p == '.') { p++; if (p >= e || *p < '0' || *p > '9') return false; while (p < e && *p >= '0' && *p <= '9') p++; }
            if (p < e && (*p == 'e' || *p == 'E')) { p++; if (p < e && (*p == '+' || *p == '-')) p++; if (p >= e || *p < '0' || *p > '9') return false; while (p < e && *p >= '0' && *p <= '9') p++; }
            v.t = JV::NUM;
            v.s.assign(s0, p - s0);
            return true;
        }
        return false;
    }
};

struct Fail { ipcfp_status st; };
[[noreturn]] void bad(ipcfp_status st = IPCFP_ERR_INVALID_ARG) { throw Fail{st}; }

const JV& need(const JV& o, const char* k, JV::T t) {
    const JV* v = o.get(k);
    if (!v || v->t != t) bad();
    return *v;
}
uint64_t u64_of(const JV& v) {   // serde: u64 fields take non-negative integer literals only
    if (v.t != JV::NUM || v.s.empty() || v.s.size() > 20) bad();
    uint64_t x = 0;
    for (char c : v.s) {
        if (c < '0' || c > '9') bad();
        uint64_t d = (uint64_t)(c - '0');
        if (x > (UINT64_MAX - d) / 10) bad();
        x = x * 10 + d;
    }
    return x;
}
int64_t i64_of(const JV& v) {   // ChainEpoch = i64
    if (v.t != JV::NUM || v.s.empty()) bad();
    bool neg = v.s[0] == '-';
    JV m;
    m.t = JV::NUM;
    m.s = neg ? v.s.substr(1) : v.s;
    uint64_t a = u64_of(m);
    if (neg) { if (a > (uint64_t)INT64_MAX + 1) bad(); return (int64_t)(0 - a); }
    if (a > (uint64_t)INT64_MAX) bad();
    return (int64_t)a;
}
void unhex(const std::string& s, std::vector<uint8_t>& out) {
    if (s.size() < 2 || s[0] != '0' || s[1] != 'x' || (s.size() & 1)) bad();
    for (size_t i = 2; i < s.size(); i += 2) {
        auto d = [&](char c) -> uint32_t { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99; };
        uint32_t h = d(s[i]), l = d(s[i + 1]);
        if (h == 99 || l == 99) bad();
        out.push_back((uint8_t)(h * 16 + l));
    }
}
// "b" + base32 lower, no padding → bytes; the C ABI carries 38-byte CIDs only
void cid_of_string(const std::string& s, uint8_t out[IPCFP_CID_LEN]) {
    if (s.empty() || s[0] != 'b') bad(IPCFP_ERR_UNSUPPORTED);
    std::vector<uint8_t> raw;
    uint32_t acc = 0;
    int bits = 0;
    for (size_t i = 1; i < s.size(); i++) {
        char c = s[i];
        uint32_t d = c >= 'a' && c <= 'z' ? (uint32_t)(c - 'a') : c >= '2' && c <= '7' ? (uint32_t)(c - '2' + 26) : 99;
        if (d == 99) bad();
        acc = (acc << 5) | d;
        bits += 5;
        if (bits >= 8) { raw.push_back((uint8_t)(acc >> (bits - 8))); bits -= 8; acc &= (1u << bits) - 1; }
    }
    if (acc != 0) bad();   // non-zero padding bits
    if (raw.size() != IPCFP_CID_LEN) bad(IPCFP_ERR_UNSUPPORTED);
    memcpy(out, raw.data(), IPCFP_CID_LEN);
}
void cid_of_field(const JV& v, uint8_t out[IPCFP_CID_LEN]) {   // ProofBlock.cid: byte array | {"/": "b…"} | "b…"
    if (v.t == JV::ARR) {
        if (v.a.size() != IPCFP_CID_LEN) bad(IPCFP_ERR_UNSUPPORTED);
        for (int k = 0; k < IPCFP_CID_LEN; k++) { uint64_t x = u64_of(v.a[k]); if (x > 255) bad(); out[k] = (uint8_t)x; }
        return;
    }
    if (v.t == JV::OBJ) { cid_of_string(need(v, "/", JV::STR).s, out); return; }
    if (v.t == JV::STR) { cid_of_string(v.s, out); return; }
    bad();
}
void unbase64(const std::string& s, std::vector<uint8_t>& out) {   // standard alphabet, padding required (base64::STANDARD)
    if (s.size() % 4) bad();
    auto d = [](char c) -> uint32_t {
        return c >= 'A' && c <= 'Z' ? (uint32_t)(c - 'A') : c >= 'a' && c <= 'z' ? (uint32_t)(c - 'a' + 26) : c >= '0' && c <= '9' ? (uint32_t)(c - '0' + 52)
               : c == '+' ? 62u : c == '/' ? 63u : 99u;
    };
    for (size_t i = 0; i < s.size(); i += 4) {
        const bool last = i + 4 == s.size();
        uint32_t a = d(s[i]), b = d(s[i + 1]);
        if (a == 99 || b == 99) bad();
        const bool p3 = s[i + 3] == '=', p2 = s[i + 2] == '=';
        if ((p2 || p3) && !last) bad();
        if (p2 && !p3) bad();
        uint32_t c = p2 ? 0 : d(s[i + 2]), e = p3 ? 0 : d(s[i + 3]);
        if (c == 99 || e == 99) bad();
        uint32_t v = (a << 18) | (b << 12) | (c << 6) | e;
        out.push_back((uint8_t)(v >> 16));
        if (!p2) out.push_back((uint8_t)(v >> 8)); else if (b & 15) bad();          // canonical: unused bits are zero
        if (!p3) out.push_back((uint8_t)v); else if (!p2 && (c & 3)) bad();
    }
}

struct Parsed {
    ipcfp_parsed_bundle pub;   // FIRST member: the handle is a pointer to it
    std::vector<uint8_t> parent_cids, child_cid, state_root;
    std::vector<ipcfp_storage_proof> sp;
    std::vector<ipcfp_event_proof> ep;
    std::vector<uint8_t> data_blob;
    std::vector<uint8_t> w_cids, w_blob;
    std::vector<uint64_t> w_offs;
    std::vector<uint32_t> w_lens;
    bool have_event_tipset = false, have_child = false, have_state_root = false;
    int64_t parent_epoch = 0, child_epoch = 0;
    bool have_child_epoch = false;
};

void same_or_set(std::vector<uint8_t>& have, bool& flag, const uint8_t* cid) {
    if (!flag) { have.assign(cid, cid + IPCFP_CID_LEN); flag = true; return; }
    if (memcmp(have.data(), cid, IPCFP_CID_LEN)) bad(IPCFP_ERR_UNSUPPORTED);
}
void child_epoch_is(Parsed& P, int64_t v) {
    if (!P.have_child_epoch) { P.child_epoch = v; P.have_child_epoch = true; return; }
    if (P.child_epoch != v) bad(IPCFP_ERR_UNSUPPORTED);
}

void read_event_proofs(Parsed& P, const JV& arr) {
    for (const JV& it : arr.a) {
        if (it.t != JV::OBJ) bad();
        const int64_t pe = i64_of(need(it, "parent_epoch", JV::NUM));
        child_epoch_is(P, i64_of(need(it, "child_epoch", JV::NUM)));
        const JV& ptc = need(it, "parent_tipset_cids", JV::ARR);
        std::vector<uint8_t> parents(ptc.a.size() * IPCFP_CID_LEN);
        for (size_t q = 0; q < ptc.a.size(); q++) { if (ptc.a[q].t != JV::STR) bad(); cid_of_string(ptc.a[q].s, parents.data() + IPCFP_CID_LEN * q); }
        if (!P.have_event_tipset) { P.parent_epoch = pe; P.parent_cids = parents; P.have_event_tipset = true; }
        else if (P.parent_epoch != pe || P.parent_cids != parents) bad(IPCFP_ERR_UNSUPPORTED);
        uint8_t c[IPCFP_CID_LEN];
        cid_of_string(need(it, "child_block_cid", JV::STR).s, c);
        same_or_set(P.child_cid, P.have_child, c);
        ipcfp_event_proof r;
        memset(&r, 0, sizeof r);
        cid_of_string(need(it, "message_cid", JV::STR).s, r.message_cid);
        r.exec_index = u64_of(need(it, "exec_index", JV::NUM));
        r.event_index = u64_of(need(it, "event_index", JV::NUM));
        const JV& ed = need(it, "event_data", JV::OBJ);
        r.emitter = u64_of(need(ed, "emitter", JV::NUM));
        const JV& tp = need(ed, "topics", JV::ARR);
        r.topics_off = P.data_blob.size();
        for (const JV& t : tp.a) {
            if (t.t != JV::STR) bad();
            const size_t before = P.data_blob.size();
            unhex(t.s, P.data_blob);
            if (P.data_blob.size() - before != 32) bad(IPCFP_ERR_UNSUPPORTED);   // the POD carries 32-byte topics (what the generator emits)
        }
        if (tp.a.size() > UINT32_MAX) bad();
        r.n_topics = (uint32_t)tp.a.size();
        r.data_off = P.data_blob.size();
        unhex(need(ed, "data", JV::STR).s, P.data_blob);
        const uint64_t dl = P.data_blob.size() - r.data_off;
        if (dl > UINT32_MAX) bad(IPCFP_ERR_UNSUPPORTED);
        r.data_len = (uint32_t)dl;
        P.ep.push_back(r);
    }
}
void read_storage_proofs(Parsed& P, const JV& arr) {
    for (const JV& it : arr.a) {
        if (it.t != JV::OBJ) bad();
        child_epoch_is(P, i64_of(need(it, "child_epoch", JV::NUM)));
        uint8_t c[IPCFP_CID_LEN];
        cid_of_string(need(it, "child_block_cid", JV::STR).s, c);
        same_or_set(P.child_cid, P.have_child, c);
        cid_of_string(need(it, "parent_state_root", JV::STR).s, c);
        same_or_set(P.state_root, P.have_state_root, c);
        ipcfp_storage_proof r;
        memset(&r, 0, sizeof r);
        r.actor_id = u64_of(need(it, "actor_id", JV::NUM));
        cid_of_string(need(it, "actor_state_cid", JV::STR).s, r.actor_state_cid);
        cid_of_string(need(it, "storage_root", JV::STR).s, r.storage_root);
        std::vector<uint8_t> b;
        unhex(need(it, "slot", JV::STR).s, b);
        if (b.size() != 32) bad();
        memcpy(r.slot, b.data(), 32);
        b.clear();
        unhex(need(it, "value", JV::STR).s, b);
        if (b.size() != 32) bad();
        memcpy(r.value, b.data(), 32);
        r.found = 1;       // the wire format carries the padded value only (storage/bundle.rs:5-14)
        r.raw_len = 32;
        P.sp.push_back(r);
    }
}
void read_blocks(Parsed& P, const JV& arr) {
    for (const JV& it : arr.a) {
        if (it.t != JV::OBJ) bad();
        const JV* c = it.get("cid");
        if (!c) bad();
        uint8_t cid[IPCFP_CID_LEN];
        cid_of_field(*c, cid);
        P.w_cids.insert(P.w_cids.end(), cid, cid + IPCFP_CID_LEN);
        const size_t off = P.w_blob.size();
        unbase64(need(it, "data", JV::STR).s, P.w_blob);
        const size_t len = P.w_blob.size() - off;
        if (len > UINT32_MAX) bad(IPCFP_ERR_UNSUPPORTED);
        P.w_offs.push_back(off);
        P.w_lens.push_back((uint32_t)len);
        while (P.w_blob.size() & 15) P.w_blob.push_back(0);   // 16-byte aligned blocks: the store's fast copy path
    }
}

}  // namespace

extern "C" {

ipcfp_status ipcfp_bundle_from_json(const char* json, uint64_t len, ipcfp_parsed_bundle** out) {
    if (!json || !out) return IPCFP_ERR_INVALID_ARG;
    *out = nullptr;
    try {
        Parser ps{json, json + len};
        JV root;
        if (!ps.value(root, 0)) return IPCFP_ERR_INVALID_ARG;
        ps.ws();
        if (ps.p != ps.e || root.t != JV::OBJ) return IPCFP_ERR_INVALID_ARG;   // trailing characters: serde_json refuses them too
        std::unique_ptr<Parsed> P(new Parsed());
        if (root.get("proofs")) read_event_proofs(*P, need(root, "proofs", JV::ARR));                       // EventProofBundle
        else {                                                                                              // UnifiedProofBundle
            read_storage_proofs(*P, need(root, "storage_proofs", JV::ARR));
            read_event_proofs(*P, need(root, "event_proofs", JV::ARR));
        }
        read_blocks(*P, need(root, "blocks", JV::ARR));
        P->w_blob.resize(P->w_blob.size() + 64, 0);
        P->data_blob.resize(P->data_blob.size() + 16, 0);
        ipcfp_parsed_bundle& b = P->pub;
        memset(&b, 0, sizeof b);
        b.tipset.parent_epoch = P->parent_epoch;
        b.tipset.child_epoch = P->child_epoch;
        b.tipset.n_parents = (uint32_t)(P->parent_cids.size() / IPCFP_CID_LEN);
        b.tipset.parent_cids = P->parent_cids.empty() ? nullptr : P->parent_cids.data();
        b.tipset.child_cid = P->have_child ? P->child_cid.data() : nullptr;
        b.tipset.child_parent_state_root = P->have_state_root ? P->state_root.data() : nullptr;
        b.n_storage_proofs = P->sp.size(); b.storage_proofs = P->sp.data();
        b.n_event_proofs = P->ep.size(); b.event_proofs = P->ep.data();
        b.data_blob = P->data_blob.data(); b.data_blob_size = P->data_blob.size() - 16;
        b.witness.n_blocks = P->w_lens.size();
        b.witness.cids = P->w_cids.data(); b.witness.offsets = P->w_offs.data(); b.witness.lengths = P->w_lens.data();
        b.witness.blob = P->w_blob.data(); b.witness.blob_size = P->w_blob.size() - 64;
        *out = &P.release()->pub;
        return IPCFP_OK;
    } catch (const Fail& f) {
        return f.st;
    } catch (const std::bad_alloc&) {
        return IPCFP_ERR_INVALID_ARG;
    }
}
void ipcfp_parsed_bundle_free(ipcfp_parsed_bundle* b) { delete reinterpret_cast<Parsed*>(b); }

}  // extern "C"
