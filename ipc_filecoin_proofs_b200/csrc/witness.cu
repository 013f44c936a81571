// witness.cu — K4/K6: turn the witness bitmap (one bit per store block, set by every recorded
// Blockstore::get) into the reference's `Vec<ProofBlock>` in `Cid` Ord order:
//   bitmap → ordered index list → radix sort by (class rank, digest) → gather CIDs + bytes.
// Replaces BTreeSet<Cid> + WitnessCollector::materialize (common/witness.rs:9-57).
#include "engine.cuh"
#include "prims.cuh"

namespace ipcfp {

struct ClassRanks { uint8_t r[IPCFP_MAX_CID_CLASSES]; };

__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}
// lexicographic order of the raw digest bytes
__device__ __forceinline__ int digest_cmp(const Digest& a, const Digest& b) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint64_t x = bswap64(a.w[k]), y = bswap64(b.w[k]);
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

__global__ void k_digest_keys(const uint32_t* __restrict__ idx, uint64_t m, const Digest* __restrict__ digests, uint32_t* keys) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint64_t w0 = digests[idx[i]].w[0];
    keys[i] = __byte_perm((uint32_t)w0, 0, 0x0123);  // first four digest bytes, big-endian
}
__global__ void k_class_keys(const uint32_t* __restrict__ idx, uint64_t m, const uint8_t* __restrict__ cls, ClassRanks cr, uint32_t* keys) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    keys[i] = cr.r[cls[idx[i]]];
}
// After the radix passes entries are ordered by (class rank, first 4 digest bytes). Runs with equal
// 4-byte prefixes (≈ m²/2³³ pairs for random digests) are finished by one thread per run.
__global__ void k_tie_fix(uint32_t* idx, uint64_t m, const Digest* __restrict__ digests, const uint8_t* __restrict__ cls, ClassRanks cr) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    auto key_eq = [&](uint64_t a, uint64_t b) {
        uint32_t x = idx[a], y = idx[b];
        return cr.r[cls[x]] == cr.r[cls[y]] && (uint32_t)digests[x].w[0] == (uint32_t)digests[y].w[0];
    };
    if (i > 0 && key_eq(i - 1, i)) return;          // not a run start
    if (i + 1 >= m || !key_eq(i, i + 1)) return;    // run of length 1
    uint64_t j = i + 1;
    while (j + 1 < m && key_eq(i, j + 1)) j++;      // run = [i, j]
    for (uint64_t a = i + 1; a <= j; a++) {         // insertion sort by full digest, stable
        uint32_t v = idx[a];
        Digest dv = digests[v];
        uint64_t b = a;
        while (b > i && digest_cmp(digests[idx[b - 1]], dv) > 0) { idx[b] = idx[b - 1]; b--; }
        idx[b] = v;
    }
}

void sort_block_indices_by_cid(Store* s, uint32_t* idx_dev, uint64_t m) {
    if (m <= 1) return;
    cudaStream_t st = s->stream;
    AsyncBuf<uint32_t> keys(m, st), keys_alt(m, st), vals_alt(m, st);
    unsigned nb = radix_blocks(m);
    AsyncBuf<uint32_t> hist((size_t)256 * nb + 256, st);
    AsyncBuf<uint64_t> scan_tmp((size_t)256 * nb + 256, st), scratch(scan_scratch_elems((uint64_t)256 * nb) + 8, st);
    ClassRanks cr{};
    for (size_t c = 0; c < s->class_rank.size(); c++) cr.r[c] = (uint8_t)s->class_rank[c];
    k_digest_keys<<<div_up(m, 256), 256, 0, st>>>(idx_dev, m, s->digests.p, keys.p); IPCFP_LAUNCH_CHECK();
    radix_sort_pairs(keys.p, idx_dev, keys_alt.p, vals_alt.p, m, 32, hist.p, scan_tmp.p, scratch.p, st);
    if (s->class_prefix.size() > 1) {
        k_class_keys<<<div_up(m, 256), 256, 0, st>>>(idx_dev, m, s->cls.p, cr, keys.p); IPCFP_LAUNCH_CHECK();
        radix_sort_pairs(keys.p, idx_dev, keys_alt.p, vals_alt.p, m, 8, hist.p, scan_tmp.p, scratch.p, st);
    }
    k_tie_fix<<<div_up(m, 256), 256, 0, st>>>(idx_dev, m, s->digests.p, s->cls.p, cr); IPCFP_LAUNCH_CHECK();
}

__global__ void k_gather_lengths(const uint32_t* __restrict__ idx, uint64_t m, const uint32_t* __restrict__ lengths, uint32_t* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = lengths[idx[i]];
}
__global__ void k_witness_cids(const uint32_t* __restrict__ idx, uint64_t m, StoreView v, uint8_t* out, const uint64_t* total, uint64_t* offsets) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) offsets[m] = *total;
    if (i >= m) return;
    uint32_t b = idx[i];
    uint8_t* o = out + 38 * i;
    uint32_t c = v.cls[b];
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = v.class_prefix[c][k];
    Digest d = v.digests[b];
#pragma unroll
    for (int w = 0; w < 4; w++)
#pragma unroll
        for (int k = 0; k < 8; k++) o[6 + 8 * w + k] = (uint8_t)(d.w[w] >> (8 * k));
}
// one warp per witness block: coalesced byte copy arena → packed witness blob
__global__ void __launch_bounds__(256) k_witness_copy(const uint32_t* __restrict__ idx, uint64_t m, StoreView v, const uint64_t* __restrict__ offsets,
                                                      uint8_t* out) {
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (w >= m) return;
    uint32_t len;
    const uint8_t* src = store_block(v, idx[w], len);
    uint8_t* dst = out + offsets[w];
    // head bytes until dst is 4-byte aligned, then word stores assembled from byte loads
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

void materialize_witness(Store* s, const uint32_t* wbits_dev, WitnessOut& out, bool to_host) {
    cudaStream_t st = s->stream;
    uint64_t n = s->n;
    uint64_t nwords = (n + 31) / 32;
    AsyncBuf<uint32_t> idx(n + 32, st);
    AsyncBuf<uint64_t> word_prefix(nwords + 8, st), scratch(scan_scratch_elems(nwords > n ? nwords : n) + 8, st);
    unsigned long long* total = s->dev_words.p + 8;
    bitmap_to_indices(wbits_dev, n, idx.p, (uint64_t*)total, word_prefix.p, scratch.p, st);
    IPCFP_CUDA(cudaMemcpyAsync(s->host_words.p + 8, total, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    uint64_t m = s->host_words.p[8];
    out.n = m;
    sort_block_indices_by_cid(s, idx.p, m);
    AsyncBuf<uint32_t> lens(m + 8, st);
    AsyncBuf<uint64_t> offs(m + 8, st);
    unsigned long long* tbytes = s->dev_words.p + 9;
    if (m) { k_gather_lengths<<<div_up(m, 256), 256, 0, st>>>(idx.p, m, s->lengths.p, lens.p); IPCFP_LAUNCH_CHECK(); }
    exclusive_scan_u32(lens.p, offs.p, m, (uint64_t*)tbytes, scratch.p, st);
    IPCFP_CUDA(cudaMemcpyAsync(s->host_words.p + 9, tbytes, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    uint64_t total_bytes = s->host_words.p[9];
    out.blob_size = total_bytes;
    AsyncBuf<uint8_t> dcids(m * 38 + 16, st), dblob(total_bytes + 16, st);
    k_witness_cids<<<div_up(m ? m : 1, 256), 256, 0, st>>>(idx.p, m, s->view, dcids.p, (const uint64_t*)tbytes, offs.p); IPCFP_LAUNCH_CHECK();
    if (m) { k_witness_copy<<<div_up(m * 32, 256), 256, 0, st>>>(idx.p, m, s->view, offs.p, dblob.p); IPCFP_LAUNCH_CHECK(); }
    if (to_host) {
        out.cids = PinnedArray(s->pool, m * 38);
        out.offsets = PinnedArray(s->pool, (m + 1) * 8);
        out.blob = PinnedArray(s->pool, total_bytes);
        if (m) IPCFP_CUDA(cudaMemcpyAsync(out.cids.p, dcids.p, m * 38, cudaMemcpyDeviceToHost, st));
        IPCFP_CUDA(cudaMemcpyAsync(out.offsets.p, offs.p, (m + 1) * 8, cudaMemcpyDeviceToHost, st));
        if (total_bytes) IPCFP_CUDA(cudaMemcpyAsync(out.blob.p, dblob.p, total_bytes, cudaMemcpyDeviceToHost, st));
    }
    out.sorted_idx = PinnedArray(s->pool, (m + 1) * 4);
    if (m) IPCFP_CUDA(cudaMemcpyAsync(out.sorted_idx.p, idx.p, m * 4, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
}

}  // namespace ipcfp
