// pass1_ring.cuh — EXPERIMENT (round 2): pass 1 with a per-lane shared-memory ring fed by cp.async.
//
// Why: k_pass1 is bound by the dependent chain "load 16-byte window → parse head → next address" — 58 % of warp time
// is long_scoreboard at the first use of a window, DRAM active 60 %, issue active 43 % (profiles/r1_ncu_full_v4_pass1.txt).
// Here every lane streams its node through a private ring of NSLOT chunks of CH bytes (chunk-aligned in the arena, so
// cp.async's 16-byte alignment holds for any block offset): chunks are requested NSLOT-1 ahead of the parser, windows
// come from shared memory (≈ 30 cycles) and L1 is bypassed (cp.async.cg). Only the fast path reads the ring; the strict
// per-event fallback, the topic comparison and nodes with links keep reading the arena.
#pragma once
#include "ipld.cuh"

namespace ipcfp {

template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CH, int NSLOT>
struct RingWin {
    static constexpr uint32_t RING = CH * NSLOT;
    static_assert((RING & (RING - 1)) == 0 && CH % 16 == 0 && NSLOT >= 2 && NSLOT <= 4, "ring geometry");
    uint32_t sbase;        // shared-window address of this lane's ring
    const uint8_t* g0;     // arena address of chunk 0 (CH-aligned, ≤ block start)
    const uint8_t* gend;   // end of the arena allocation: nothing is read at or past it
    uint32_t skew;         // block start − g0: node offset x lives at ring offset (x + skew) mod RING
    uint32_t nchunks;      // chunks covering the node plus the 24-byte over-read of a window
    uint32_t issued;       // chunks requested so far (chunk k → slot k mod NSLOT)
    uint32_t done;         // chunks known to have landed

    __device__ __forceinline__ void init(uint32_t sbase_, const uint8_t* p, uint32_t len, const uint8_t* gend_) {
        sbase = sbase_;
        g0 = (const uint8_t*)((uintptr_t)p & ~(uintptr_t)(CH - 1));
        gend = gend_;
        skew = (uint32_t)(p - g0);
        nchunks = (skew + len + 32 + CH - 1) / CH;
        issued = done = 0;
    }
    __device__ __forceinline__ void issue_one() {
        const uint8_t* src = g0 + (size_t)issued * CH;
        const uint32_t dst = sbase + (issued % NSLOT) * CH;
        if (src + CH <= gend) {
#pragma unroll
            for (int k = 0; k < CH / 16; k++) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(src + 16 * k) : "memory");
        } else {   // last chunk of the arena: read only what exists, zero-fill the rest
#pragma unroll
            for (int k = 0; k < CH / 16; k++) {
                const uint8_t* s = src + 16 * k;
                long long left = gend - s;
                uint32_t sz = left >= 16 ? 16u : (left > 0 ? (uint32_t)left : 0u);
                if (sz == 0) s = gend - 16;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16 * k), "l"(s), "r"(sz) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        issued++;
    }
    // slots of chunks below lo_chunk are free: keep NSLOT chunks requested from there on
    __device__ __forceinline__ void top_up(uint32_t lo_chunk) {
        while (issued < nchunks && issued < lo_chunk + NSLOT) issue_one();
    }
    // chunks 0..hi_chunk must have landed before they are read
    __device__ __forceinline__ void need(uint32_t hi_chunk) {
        if (hi_chunk < done) return;
        uint32_t later = issued - 1 - hi_chunk;          // requests made after hi_chunk may stay in flight
        if (later >= 3) { cp_async_wait<3>(); done = issued - 3; }
        else if (later == 2) { cp_async_wait<2>(); done = issued - 2; }
        else if (later == 1) { cp_async_wait<1>(); done = issued - 1; }
        else { cp_async_wait<0>(); done = issued; }
    }
    __device__ __forceinline__ uint2 lds(uint32_t off) const {
        uint2 v;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(sbase + (off & (RING - 1))));
        return v;
    }
    // 16 bytes of the node at offset pos (same contract as win_load)
    __device__ __forceinline__ void load(uint32_t pos, uint64_t& w0, uint64_t& w1) {
        const uint32_t a = skew + pos, a0 = a & ~7u;
        top_up(a0 / CH);
        need((a0 + 23) / CH);
        const uint32_t s = (a & 7) * 8;
        const uint2 x0 = lds(a0), x1 = lds(a0 + 8), x2 = lds(a0 + 16);
        const bool up = (s & 32) != 0;
        const uint32_t c0 = up ? x0.y : x0.x, c1 = up ? x1.x : x0.y, c2 = up ? x1.y : x1.x, c3 = up ? x2.x : x1.y, c4 = up ? x2.y : x2.x;
        w0 = (uint64_t)__funnelshift_r(c0, c1, s) | ((uint64_t)__funnelshift_r(c1, c2, s) << 32);
        w1 = (uint64_t)__funnelshift_r(c2, c3, s) | ((uint64_t)__funnelshift_r(c3, c4, s) << 32);
    }
    // generic pointer to node offset 0 for byte-wise reads of the first `upto` bytes (must not wrap: skew + upto ≤ RING)
    __device__ __forceinline__ const uint8_t* head_ptr(uint32_t upto) {
        top_up(0);
        const uint32_t hi = (skew + upto) / CH;
        need(hi < nchunks ? hi : nchunks - 1);
        return (const uint8_t*)__cvta_shared_to_generic((size_t)sbase) + skew;
    }
};

}  // namespace ipcfp
