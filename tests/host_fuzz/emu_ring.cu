// emu_ring.cu — EXPERIMENT SUPPORT (round 2): the shared-memory ring variant of pass 1 (csrc/pass1_ring.cuh) on the CPU.
//
// `pass1_ring_item` is run over generated and mutated events-AMT root blocks at random arena positions (including the very end of
// the arena, where cp.async must zero-fill) with the adversarial copy model of RingWin's host build (a request poisons its slot and
// delivers only when a wait lets it complete), for the three ring geometries, against the arena decode sequence of pass1_body.
// Property: whenever the ring path takes a node, the arena path accepts it too with the same (any, #proofs, #bytes); and every
// well-formed single-node block IS taken.
//
//   nvcc -std=c++17 -O2 -o emu_ring tests/host_fuzz/emu_ring.cu && ./emu_ring [iterations] [seed]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "host_shims.h"

#include "../../ipc_filecoin_proofs_b200/csrc/ipld.cuh"
#ifndef __CUDA_ARCH__
#define prefetch_l2(p) ((void)0)
#define prefetch_l1(p) ((void)0)
#endif
#include "../../ipc_filecoin_proofs_b200/csrc/pass1_ring.cuh"

using namespace ipcfp;

static uint64_t rng_state;
static uint64_t rnd() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void put_head(std::vector<uint8_t>& o, int major, uint64_t v) {
    if (v < 24) o.push_back((uint8_t)(major << 5 | v));
    else if (v < 0x100) { o.push_back((uint8_t)(major << 5 | 24)); o.push_back((uint8_t)v); }
    else if (v < 0x10000) { o.push_back((uint8_t)(major << 5 | 25)); o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)v); }
    else if (v < 0x100000000ull) { o.push_back((uint8_t)(major << 5 | 26)); for (int s = 24; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }
    else { o.push_back((uint8_t)(major << 5 | 27)); for (int s = 56; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }
}
static uint8_t T0[32], T1[32];
static void put_entry(std::vector<uint8_t>& o, uint64_t flags, const char* key, uint64_t codec, size_t vlen, const uint8_t* fixed = nullptr) {
    put_head(o, 4, 4);
    put_head(o, 0, flags);
    put_head(o, 3, strlen(key));
    o.insert(o.end(), key, key + strlen(key));
    put_head(o, 0, codec);
    put_head(o, 2, vlen);
    for (size_t i = 0; i < vlen; i++) o.push_back(fixed ? fixed[i] : (uint8_t)rnd());
}
static void make_event(std::vector<uint8_t>& o) {
    put_head(o, 4, 2);
    static const uint64_t emitters[] = {5, 24, 255, 1001, 1001, 65536, (1ull << 40) + 1001};
    put_head(o, 0, emitters[rnd() % 7]);
    unsigned shape = (unsigned)(rnd() % 8);
    bool hit = rnd() % 6 == 0;
    if (shape == 0) {
        put_head(o, 4, 2);
        uint8_t tp[64]; memcpy(tp, T0, 32); memcpy(tp + 32, T1, 32);
        put_entry(o, 3, "topics", 0x55, hit ? 64 : 32 * (rnd() % 5), hit ? tp : nullptr);
        put_entry(o, 3, "data", 0x55, rnd() % 300);
    } else {
        unsigned nt = 1 + (unsigned)(rnd() % 4);
        if (hit && nt < 2) nt = 2;
        bool has_d = rnd() % 4 != 0;
        put_head(o, 4, nt + (has_d ? 1 : 0));
        static const char* tk[] = {"t1", "t2", "t3", "t4"};
        for (unsigned t = 0; t < nt; t++) put_entry(o, 3, tk[t], rnd() % 16 == 0 ? rnd() % 24 : 0x55, rnd() % 32 == 0 ? rnd() % 40 : 32, hit && t < 2 ? (t ? T1 : T0) : nullptr);
        if (has_d) put_entry(o, 3, "d", 0x55, rnd() % 6 == 0 ? 256 + rnd() % 700 : rnd() % 64);
    }
}
static std::vector<uint8_t> make_root(bool& wellformed_single) {
    std::vector<uint8_t> o;
    uint32_t bw = rnd() % 2 ? 5u : 3u, width = 1u << bw, nmax = width < 20 ? width : 20;
    uint32_t n = (uint32_t)(rnd() % (nmax + 1));
    std::vector<uint8_t> bm(bw <= 3 ? 1 : (1u << (bw - 3)), 0);
    for (uint32_t k = 0; k < n;) { uint32_t b = (uint32_t)(rnd() % width); if (!(bm[b / 8] >> (b % 8) & 1)) { bm[b / 8] |= (uint8_t)(1u << (b % 8)); k++; } }
    put_head(o, 4, 4); put_head(o, 0, bw); put_head(o, 0, 0); put_head(o, 0, n);
    put_head(o, 4, 3); put_head(o, 2, bm.size()); o.insert(o.end(), bm.begin(), bm.end());
    put_head(o, 4, 0); put_head(o, 4, n);
    for (uint32_t k = 0; k < n; k++) make_event(o);
    wellformed_single = true;
    return o;
}

template <int CH, int NSLOT>
static int run(uint64_t iters, uint64_t* taken_out, uint64_t* wf_out) {
    std::vector<uint8_t> arena, ringbuf(CH * NSLOT + 16);
    uint64_t taken_n = 0, wf = 0;
    Matcher m;
    memset(&m, 0, sizeof m);
    memcpy(m.t0, T0, 32); memcpy(m.t1, T1, 32);
    m.actor = 1001; m.has_actor = rnd() % 2;
    for (uint64_t it = 0; it < iters; it++) {
        bool wfs;
        std::vector<uint8_t> blk = make_root(wfs);
        unsigned nmut = it % 3 == 0 ? 1 + (unsigned)(rnd() % 2) : 0;
        for (unsigned k = 0; k < nmut; k++) {
            size_t at = rnd() % blk.size();
            switch (rnd() % 4) {
                case 0: blk[at] = (uint8_t)rnd(); break;
                case 1: blk[at] ^= (uint8_t)(1u << (rnd() % 8)); break;
                case 2: blk.erase(blk.begin() + (long)at); break;
                default: blk.insert(blk.begin() + (long)at, (uint8_t)rnd()); break;
            }
            if (blk.empty()) blk.push_back(0x84);
            wfs = false;
        }
        // arena: [16 pad][junk lead][block][junk tail or nothing][32 pad]; the allocation base is 256-byte aligned like cudaMalloc's
        unsigned lead = (unsigned)(rnd() % 700);
        bool at_end = rnd() % 4 == 0;
        unsigned tail = at_end ? 0 : (unsigned)(rnd() % 900);
        size_t total = 16 + lead + blk.size() + tail + 32;
        arena.assign(total + 256, 0);
        uint8_t* base = (uint8_t*)(((uintptr_t)arena.data() + 255) & ~(uintptr_t)255);
        for (size_t k = 0; k < 16 + lead; k++) base[k] = (uint8_t)rnd();
        memcpy(base + 16 + lead, blk.data(), blk.size());
        for (size_t k = 0; k < tail; k++) base[16 + lead + blk.size() + k] = (uint8_t)rnd();
        const uint8_t* p = base + 16 + lead;
        const uint32_t len = (uint32_t)blk.size();
        const uint8_t* arena_end = base + total;
        // arena path (pass1_body's sequence)
        Rd r(p, len);
        uint32_t bw, height;
        uint64_t cnt;
        amt_root_begin(r, 3, bw, height, cnt);
        AmtNodeHdr h;
        amt_node_begin(r, bw, h);
        uint32_t nv = rd_array(r);
        WalkOut wa{0, 0, false};
        node_events<WALK_COUNT>(r, p, h, nv, 0, m, wa, nullptr, 0);
        amt_node_finish(r, h, nv, height);
        // ring path
        RingWin<CH, NSLOT> ring;
        for (auto& b : ringbuf) b = 0xAB;
        ring.init(ringbuf.data(), p, len, arena_end);
        ring.top_up(0);
        WalkOut wr{0, 0, false};
        bool taken = pass1_ring_item(ring, p, len, m, wr);
        if (taken) {
            taken_n++;
            if (r.err || h.nl || wr.any != wa.any || wr.nproofs != wa.nproofs || wr.nbytes != wa.nbytes) {
                fprintf(stderr, "RING MISMATCH <%d,%d> at iteration %llu: ring took the node (any %d np %u nb %u), arena err %u nl %u (any %d np %u nb %u); lead %u len %u at_end %d\n", CH, NSLOT,
                        (unsigned long long)it, wr.any, wr.nproofs, wr.nbytes, r.err, h.nl, wa.any, wa.nproofs, wa.nbytes, lead, len, at_end);
                return 1;
            }
        }
        if (wfs) {
            wf++;
            if (!taken) { fprintf(stderr, "RING <%d,%d>: a well-formed single-node block was not taken (iteration %llu, len %u, lead %u)\n", CH, NSLOT, (unsigned long long)it, len, lead); return 1; }
        }
    }
    *taken_out += taken_n; *wf_out += wf;
    return 0;
}

int main(int argc, char** argv) {
    uint64_t iters = argc > 1 ? strtoull(argv[1], nullptr, 10) : 200000;
    rng_state = argc > 2 ? strtoull(argv[2], nullptr, 10) : 0xC0FFEEull;
    for (int i = 0; i < 32; i++) { T0[i] = (uint8_t)rnd(); T1[i] = (uint8_t)rnd(); }
    uint64_t taken = 0, wf = 0;
    if (run<128, 2>(iters, &taken, &wf) || run<128, 4>(iters, &taken, &wf) || run<256, 2>(iters, &taken, &wf)) return 1;
    printf("ok: ring pass 1 == arena pass 1 for 3 geometries x %llu blocks: %llu nodes taken by the ring path, %llu well-formed single-node blocks (all taken)\n",
           (unsigned long long)iters, (unsigned long long)taken, (unsigned long long)wf);
    return 0;
}
