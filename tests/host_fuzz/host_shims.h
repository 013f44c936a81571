// host_shims.h — lets the product's device headers compile for the HOST (nvcc host pass): every __device__ function becomes
// __host__ __device__ and the handful of intrinsics they use map to compiler builtins. TEST INFRASTRUCTURE ONLY.
// Include BEFORE any csrc/*.cuh header.
#pragma once
#undef __device__
#define __device__ __location__(host) __location__(device)
#ifndef __CUDA_ARCH__
static inline unsigned host_funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
#define __funnelshift_r(lo, hi, s) host_funnelshift_r((lo), (hi), (s))
static inline unsigned host_byte_perm(unsigned x, unsigned y, unsigned s) {
    unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
#define __byte_perm(x, y, s) host_byte_perm((x), (y), (s))
#define __popc(x) __builtin_popcount(x)
#define __popcll(x) __builtin_popcountll(x)
#define __ffs(x) __builtin_ffs(x)
#define __ffsll(x) __builtin_ffsll(x)
#define __ldg(p) (*(p))
#define atomicMin(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define atomicOr(p, v) (*(p) |= (v))
#endif

