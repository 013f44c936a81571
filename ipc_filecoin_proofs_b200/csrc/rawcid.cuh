// rawcid.cuh — a message CID as five aligned words (values of the BLS/SECP message AMTs).
#pragma once
#include "common.cuh"

namespace ipcfp {

// w[0..3] = digest bytes (memory order), w[4] low 48 bits = CID prefix bytes 0..5
struct RawCid { uint64_t w[5]; };

#ifdef __CUDACC__
__device__ __forceinline__ bool rawcid_eq(const RawCid& x, const RawCid& y) {
    return x.w[0] == y.w[0] && x.w[1] == y.w[1] && x.w[2] == y.w[2] && x.w[3] == y.w[3] && x.w[4] == y.w[4];
}
__device__ __forceinline__ uint64_t rawcid_hash(const RawCid& c) { return mix64(c.w[0] ^ (c.w[2] * 0x9E3779B97F4A7C15ULL) ^ c.w[4]); }
#endif

}  // namespace ipcfp
