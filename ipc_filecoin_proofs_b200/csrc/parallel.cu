// parallel.cu — device helpers for the cross-shard part of the path (one process per GPU; the
// collectives themselves belong to the caller: torch.distributed / NCCL).
//
// The reference builds the execution order by walking every message AMT and dropping repeated CIDs
// "first seen wins" (events/utils.rs:56-91). With the message list sharded over ranks, duplicates can
// span shards, so the dedup is a distributed hash join:
//   ipcfp_exec_bucketize   every rank routes (cid, global position) of its slice to owner = hash(cid) % world
//   [all-to-all]
//   ipcfp_exec_dedup       every owner finds, per distinct CID, the smallest position; all other positions
//                          of that CID are duplicates → returned as a list
//   [all-gather of the (tiny) duplicate lists]  →  exec index i ↔ raw position p(i) on the host
//   ipcfp_exec_fetch       CIDs at requested raw positions (for EventProof.message_cid)
#include <algorithm>
#include <vector>

#include "engine.cuh"
#include "prims.cuh"
#include "rawcid.cuh"

namespace ipcfp {

struct ExecEntry { RawCid c; uint64_t pos; };  // 48 bytes

// owner of every record of the slice (key for the stable partition)
__global__ void k_exec_owner(const RawCid* __restrict__ seg, uint64_t nseg, uint32_t world, uint32_t* keys, uint32_t* vals) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nseg) return;
    keys[i] = (uint32_t)((rawcid_hash(seg[i]) >> 32) % world);
    vals[i] = (uint32_t)i;
}
// first sorted position of every owner
__global__ void k_exec_starts(const uint32_t* __restrict__ keys, uint64_t nseg, unsigned long long* start) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nseg) return;
    if (j == 0 || keys[j - 1] != keys[j]) start[keys[j]] = j;
}
// entries leave in (owner, position) order: within a bucket the global positions are increasing
__global__ void k_exec_scatter(const RawCid* __restrict__ seg, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t nseg,
                               uint64_t pos0, const unsigned long long* __restrict__ start, uint64_t cap, ExecEntry* send) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nseg) return;
    uint32_t owner = keys[j], src = vals[j];
    uint64_t slot = j - start[owner];
    if (slot < cap) { ExecEntry e; e.c = seg[src]; e.pos = pos0 + src; send[(uint64_t)owner * cap + slot] = e; }
}

// entry k of the received buffer (world segments of `cap`, counts[r] valid in segment r)
__device__ __forceinline__ const ExecEntry* recv_entry(const ExecEntry* recv, const uint64_t* seg_off, uint32_t world, uint64_t cap, uint64_t k) {
    uint32_t r = 0;
    while (r + 1 < world && k >= seg_off[r + 1]) r++;
    return recv + (uint64_t)r * cap + (k - seg_off[r]);
}
// pass 1: one canonical slot per distinct CID (value = fingerprint << 32 | smallest entry ordinal + 1)
__global__ void k_exec_claim(const ExecEntry* __restrict__ recv, const uint64_t* __restrict__ seg_off, uint32_t world, uint64_t cap, uint64_t total,
                             unsigned long long* table, uint64_t mask) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const ExecEntry* e = recv_entry(recv, seg_off, world, cap, k);
    uint64_t h = rawcid_hash(e->c);
    uint32_t fp = (uint32_t)(h >> 40) | 1u;   // bits 40..63: independent of the slot bits
    unsigned long long mine = ((unsigned long long)fp << 32) | (unsigned long long)(k + 1);
    uint64_t slot = h & mask;   // low bits: independent of the owner choice ((h >> 32) % world)
    for (;;) {
        unsigned long long v = table[slot];
        if (v == 0) { v = atomicCAS(&table[slot], 0ull, mine); if (v == 0) return; }
        if ((uint32_t)(v >> 32) == fp && rawcid_eq(recv_entry(recv, seg_off, world, cap, (uint32_t)v - 1)->c, e->c)) { atomicMin(&table[slot], mine); return; }
        slot = (slot + 1) & mask;
    }
}
__device__ __forceinline__ uint64_t exec_find_slot(const ExecEntry* recv, const uint64_t* seg_off, uint32_t world, uint64_t cap, const ExecEntry* e,
                                                   const unsigned long long* table, uint64_t mask) {
    uint64_t h = rawcid_hash(e->c);
    uint32_t fp = (uint32_t)(h >> 40) | 1u;   // bits 40..63: independent of the slot bits
    uint64_t slot = h & mask;   // low bits: independent of the owner choice ((h >> 32) % world)
    for (;;) {
        unsigned long long v = table[slot];
        if ((uint32_t)(v >> 32) == fp && rawcid_eq(recv_entry(recv, seg_off, world, cap, (uint32_t)v - 1)->c, e->c)) return slot;
        slot = (slot + 1) & mask;
    }
}
// pass 2: the received buffer is ordered by (sender rank, position), so the smallest entry ordinal of a CID is
// its smallest global position: every other entry of that CID is a duplicate
__global__ void k_exec_dups(const ExecEntry* __restrict__ recv, const uint64_t* __restrict__ seg_off, uint32_t world, uint64_t cap, uint64_t total,
                            const unsigned long long* table, uint64_t mask, uint64_t* dup, uint64_t cap_out, unsigned long long* n_dup) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const ExecEntry* e = recv_entry(recv, seg_off, world, cap, k);
    uint64_t slot = exec_find_slot(recv, seg_off, world, cap, e, table, mask);
    if ((uint32_t)table[slot] - 1 != (uint32_t)k) {
        unsigned long long j = atomicAdd(n_dup, 1ull);
        if (j < cap_out) dup[j] = e->pos;
    }
}
__global__ void k_exec_fetch(const RawCid* __restrict__ seg, uint64_t nseg, uint64_t pos0, const uint64_t* __restrict__ req, uint64_t n, RawCid* out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t p = req[j];
    if (p >= pos0 && p - pos0 < nseg) out[j] = seg[p - pos0];
}

// grow-only scratch + a private non-blocking stream per device for the helpers (no allocation in steady state)
struct HelperCtx {
    cudaStream_t st = nullptr;
    DevBuf<unsigned long long> table, minpos, small;
};
static HelperCtx& helper_ctx(int device) {
    static HelperCtx ctx[16];
    HelperCtx& c = ctx[device & 15];
    if (!c.st) { IPCFP_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking)); c.small.alloc(1024); }
    return c;
}

void exec_bucketize(int device, const void* seg, uint64_t nseg, uint64_t pos0, uint32_t world, uint64_t cap, void* send, uint64_t* counts_host) {
    check_device(device);
    if (!world || world > 256) throw Error(IPCFP_ERR_INVALID_ARG, "bad world size");
    if (nseg >= 0xffffffffull) throw Error(IPCFP_ERR_UNSUPPORTED, "slice too long");
    HelperCtx& hc = helper_ctx(device);
    cudaStream_t st = hc.st;
    for (uint32_t r = 0; r < world; r++) counts_host[r] = 0;
    if (!nseg) return;
    unsigned nb = radix_blocks(nseg);
    AsyncBuf<uint32_t> keys(nseg, st), vals(nseg, st), ka(nseg, st), va(nseg, st), hist((size_t)256 * nb + 256, st);
    AsyncBuf<uint64_t> scan_tmp((size_t)256 * nb + 256, st), scratch(scan_scratch_elems((uint64_t)256 * nb) + 8, st);
    unsigned long long* start = hc.small.p;   // [0, world]
    std::vector<unsigned long long> h_start(world + 1, nseg);
    IPCFP_CUDA(cudaMemcpyAsync(start, h_start.data(), (world + 1) * 8, cudaMemcpyHostToDevice, st));
    k_exec_owner<<<div_up(nseg, 256), 256, 0, st>>>((const RawCid*)seg, nseg, world, keys.p, vals.p); IPCFP_LAUNCH_CHECK();
    radix_sort_pairs(keys.p, vals.p, ka.p, va.p, nseg, 8, hist.p, scan_tmp.p, scratch.p, st);
    k_exec_starts<<<div_up(nseg, 256), 256, 0, st>>>(keys.p, nseg, start); IPCFP_LAUNCH_CHECK();
    k_exec_scatter<<<div_up(nseg, 256), 256, 0, st>>>((const RawCid*)seg, keys.p, vals.p, nseg, pos0, start, cap, (ExecEntry*)send); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaMemcpyAsync(h_start.data(), start, (world + 1) * 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    // owners without entries keep start == nseg; counts from consecutive starts of present owners
    uint64_t next = nseg;
    for (int r = (int)world - 1; r >= 0; r--) {
        if (h_start[r] == nseg) { counts_host[r] = 0; continue; }
        counts_host[r] = next - h_start[r];
        next = h_start[r];
    }
    for (uint32_t r = 0; r < world; r++) if (counts_host[r] > cap) throw Error(IPCFP_ERR_INVALID_ARG, "bucket capacity too small", counts_host[r]);
}
void exec_dedup(int device, const void* recv, const uint64_t* counts, uint32_t world, uint64_t cap, uint64_t* dup_dev, uint64_t cap_out, uint64_t* n_dup) {
    check_device(device);
    HelperCtx& hc = helper_ctx(device);
    cudaStream_t st = hc.st;
    std::vector<uint64_t> seg(world + 1, 0);
    for (uint32_t r = 0; r < world; r++) { if (counts[r] > cap) throw Error(IPCFP_ERR_INVALID_ARG, "count exceeds bucket capacity"); seg[r + 1] = seg[r] + counts[r]; }
    uint64_t total = seg[world];
    *n_dup = 0;
    if (!total) return;
    if (total >= 0xffffffffull) throw Error(IPCFP_ERR_UNSUPPORTED, "more than 2^32 messages per owner");
    uint64_t slots = 64;
    while (slots < 2 * total) slots <<= 1;
    if (world + 2 > 1000) throw Error(IPCFP_ERR_UNSUPPORTED, "world too large");
    hc.table.ensure(slots);
    uint64_t* d_seg = (uint64_t*)hc.small.p;             // [0, world]
    unsigned long long* nd = hc.small.p + 1023;
    IPCFP_CUDA(cudaMemsetAsync(hc.table.p, 0, slots * 8, st));
    IPCFP_CUDA(cudaMemsetAsync(nd, 0, 8, st));
    IPCFP_CUDA(cudaMemcpyAsync(d_seg, seg.data(), (world + 1) * 8, cudaMemcpyHostToDevice, st));
    unsigned g = div_up(total, 256);
    const ExecEntry* e = (const ExecEntry*)recv;
    k_exec_claim<<<g, 256, 0, st>>>(e, d_seg, world, cap, total, hc.table.p, slots - 1); IPCFP_LAUNCH_CHECK();
    k_exec_dups<<<g, 256, 0, st>>>(e, d_seg, world, cap, total, hc.table.p, slots - 1, dup_dev, cap_out, nd); IPCFP_LAUNCH_CHECK();
    unsigned long long n = 0;
    IPCFP_CUDA(cudaMemcpyAsync(&n, nd, 8, cudaMemcpyDeviceToHost, st));
    IPCFP_CUDA(cudaStreamSynchronize(st));
    if (n > cap_out) throw Error(IPCFP_ERR_INVALID_ARG, "duplicate list capacity too small", n);
    *n_dup = n;
}
void exec_fetch(int device, const void* seg, uint64_t nseg, uint64_t pos0, const uint64_t* req_dev, uint64_t n, void* out_dev) {
    check_device(device);
    if (!n) return;
    k_exec_fetch<<<div_up(n, 256), 256>>>((const RawCid*)seg, nseg, pos0, req_dev, n, (RawCid*)out_dev); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaStreamSynchronize(nullptr));
}

}  // namespace ipcfp

// =====================================================================================================================
// In-library cross-shard protocol over NCCL (ipcfp_comm_init / ipcfp_generate_event_proof_sharded).
//
// One process per GPU. Receipts shard by index range (events/generator.rs:209-301 is independent per receipt); what spans
// shards is (1) the execution order — every message AMT concatenated, first occurrence of a CID wins (events/utils.rs:48-94) —
// and (2) the union of the per-shard witness CID sets (common/witness.rs:24-40). Everything is enqueued on CUDA streams with
// sizes the host already knows; the host waits for its peers twice (H0: does every shard have a message list, how long; H2:
// did every shard get through pass 2, how many matches / witness blocks), both while its own GPU is busy:
//
//   H0   all-gather  {ok, Nraw, nseg}                                      → global positions, exact buffer sizes, common abort
//   X    bucketize by hash(cid) % world → all-to-all (ncclSend/ncclRecv group) → first-seen dedup on the owners → the owners set
//        one bit per NON-first occurrence in a bitmap over the raw positions → all-reduce (sum of disjoint bitmaps = OR)
//        [exchange stream: runs underneath pass 1]
//   P    n_exec = zero bits; exec index i ↔ position of the (i+1)-th zero bit (prefix popcount + select) for the rank's matches;
//        pass 2 runs with the global n_exec (so "Missing message at index" keeps its place in the error order)
//   H2   all-gather  {first error key, matches, proofs, witness blocks}    → all ranks fail together with the SAME error
//   F    all-gather of the wanted positions → every owner copies the CIDs it holds → all-reduce → EventProof.message_cid patched
//   W    all-gather of the sorted per-shard witness CID lists → bucketed k-way merge + unique on every rank
// NCCL is resolved with dlopen at ipcfp_comm_init (libnccl.so.2: the copy already in the process — e.g. PyTorch's — or the
// system one), so the library itself keeps linking only cudart and loads on machines without NCCL.
// =====================================================================================================================
#include <cstdio>
#include <dlfcn.h>
#include <nccl.h>

namespace ipcfp {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GetVersion)(int*);
};
static NcclApi* nccl_api() {
    static std::mutex mu;
    static NcclApi api;
    static bool ready = false;
    std::lock_guard<std::mutex> g(mu);
    if (ready) return &api;
    const char* names[] = {getenv("IPCFP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break;
    if (!h) throw Error(IPCFP_ERR_NCCL, std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""));
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) throw Error(IPCFP_ERR_NCCL, std::string("NCCL symbol missing: ") + n); return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    ready = true;
    return &api;
}
#define IPCFP_NCCL(expr)                                                                                                     \
    do {                                                                                                                     \
        ncclResult_t _r = (expr);                                                                                            \
        if (_r != ncclSuccess) throw ::ipcfp::Error(IPCFP_ERR_NCCL, std::string(#expr) + ": " + nccl_api()->GetErrorString(_r)); \
    } while (0)

struct Comm {
    int device = 0;
    uint32_t world = 1, rank = 0;
    ncclComm_t cx = nullptr, cw = nullptr;   // exchange (execution order) / witness union: independent streams, independent communicators
    cudaStream_t sx = nullptr, sw = nullptr;   // exchange / fetch stream (communicator cx); witness-union stream (communicator cw)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    cudaEvent_t tm[10] = {};                 // timing: exchange begin/end (sx), fetch begin/end, union begin/end (engine stream); [6..8] inside the exchange: bucketize | all-to-all | dedup done
    // grow-only device scratch (allocated during warm-up, then reused)
    DevBuf<uint8_t> sendbuf, recvbuf, gather, merged, recs, part_send;
    DevBuf<unsigned long long> table, words, words2;
    DevBuf<uint32_t> bitmap, bitmap_sum, zeros, flags, starts;
    DevBuf<uint64_t> zprefix, scan_tmp, req, req_pad, req_all, ans, ans_sum, fscan;
    DevBuf<uint32_t> pos_of;
    DevBuf<unsigned long long> part_words;   // partitioned union: [0, W] piece bounds | [W+1, 2W+1) received piece lengths | [n_part, overflow] | the same of all ranks
    PinnedBuf<uint64_t> host;                // mapped: H0 / H2 read-backs
    ~Comm() {
        cudaSetDevice(device);
        NcclApi* n = nullptr;
        try { n = nccl_api(); } catch (...) {}
        if (n) { if (cx) n->CommDestroy(cx); if (cw) n->CommDestroy(cw); }
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (ev_c) cudaEventDestroy(ev_c);
        for (auto& e : tm) if (e) cudaEventDestroy(e);
        if (sx) cudaStreamDestroy(sx);
        if (sw) cudaStreamDestroy(sw);
    }
};

void comm_unique_id(uint8_t* id128) {
    ncclUniqueId id;
    IPCFP_NCCL(nccl_api()->GetUniqueId(&id));
    static_assert(sizeof id == IPCFP_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id128, &id, sizeof id);
}
Comm* comm_init(const uint8_t* id128, uint32_t world, uint32_t rank, int device) {
    check_device(device);
    if (!world || world > 256 || rank >= world) throw Error(IPCFP_ERR_INVALID_ARG, "bad world size / rank");
    NcclApi* n = nccl_api();
    std::unique_ptr<Comm> c(new Comm());
    c->device = device; c->world = world; c->rank = rank;
    {   // the exchange must not queue behind the full-grid scan kernels of the engine stream
        int lo_p = 0, hi_p = 0;
        IPCFP_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        IPCFP_CUDA(cudaStreamCreateWithPriority(&c->sx, cudaStreamNonBlocking, hi_p));
        IPCFP_CUDA(cudaStreamCreateWithPriority(&c->sw, cudaStreamNonBlocking, hi_p));
    }
    IPCFP_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
    IPCFP_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
    IPCFP_CUDA(cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming));
    for (auto& e : c->tm) IPCFP_CUDA(cudaEventCreate(&e));
    c->host.alloc(4096);    // mapped: [0, 2048) gathered words of H0 / H2
    c->words.alloc(4096);
    c->words2.alloc(4096);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    IPCFP_NCCL(n->CommInitRank(&c->cx, (int)world, id, (int)rank));
    // the second communicator's id travels over the first one
    ncclUniqueId id2;
    if (rank == 0) IPCFP_NCCL(n->GetUniqueId(&id2));
    IPCFP_CUDA(cudaMemcpyAsync(c->words.p, &id2, sizeof id2, cudaMemcpyHostToDevice, c->sx));
    IPCFP_NCCL(n->Broadcast(c->words.p, c->words.p, sizeof id2, ncclUint8, 0, c->cx, c->sx));
    IPCFP_CUDA(cudaMemcpyAsync(&id2, c->words.p, sizeof id2, cudaMemcpyDeviceToHost, c->sx));
    IPCFP_CUDA(cudaStreamSynchronize(c->sx));
    IPCFP_NCCL(n->CommInitRank(&c->cw, (int)world, id2, (int)rank));
    return c.release();
}
void comm_destroy(Comm* c) { delete c; }
uint32_t comm_world(const Comm* c) { return c->world; }
uint32_t comm_rank(const Comm* c) { return c->rank; }

// ------------------------------------------------------------------------------------------ exchange kernels
#define XSEG_HDR 48   // a segment = [count u64, 40 bytes pad][cap entries of 48 bytes]

// counts per owner from the first sorted position of every owner (nseg where an owner has no entry); writes the segment headers
__global__ void k_exec_seg_headers(const unsigned long long* __restrict__ start, uint64_t nseg, uint32_t world, uint64_t cap, uint8_t* send,
                                   unsigned long long* overflow) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long next = nseg;
    for (int r = (int)world - 1; r >= 0; r--) {
        unsigned long long cnt = 0;
        if (start[r] != nseg) { cnt = next - start[r]; next = start[r]; }
        if (cnt > cap) { *overflow = 1; cnt = cap; }
        *(unsigned long long*)(send + (uint64_t)r * (XSEG_HDR + cap * 48)) = cnt;
    }
}
__global__ void k_exec_scatter_seg(const RawCid* __restrict__ seg, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t nseg,
                                   uint64_t pos0, const unsigned long long* __restrict__ start, uint64_t cap, uint8_t* send) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nseg) return;
    uint32_t owner = keys[j], src = vals[j];
    uint64_t slot = j - start[owner];
    if (slot < cap) {
        ExecEntry e; e.c = seg[src]; e.pos = pos0 + src;
        *(ExecEntry*)(send + (uint64_t)owner * (XSEG_HDR + cap * 48) + XSEG_HDR + slot * 48) = e;
    }
}
// ---- order-preserving partition by owner in three kernels (count per warp run → one scan → scatter), no key / value arrays ----
#define XB_RUN 256u   // consecutive entries one warp handles (8 chunks of 32)
__device__ __forceinline__ uint32_t exec_owner_of(const RawCid& c, uint32_t world) { return (uint32_t)((rawcid_hash(c) >> 32) % world); }
// cnt[owner * nruns + run] = entries of that owner in run `run` of the slice
__global__ void __launch_bounds__(128) k_xb_count(const RawCid* __restrict__ seg, uint64_t nseg, uint32_t world, uint32_t nruns, uint32_t* cnt) {
    const uint32_t run = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (run >= nruns) return;
    const uint64_t base = (uint64_t)run * XB_RUN;
    uint32_t mine = 0;                                   // lane o (< 32) accumulates owner o; world > 32: lanes take owners o, o+32, … in turn
    for (uint32_t c = 0; c < XB_RUN / 32; c++) {
        const uint64_t i = base + c * 32 + lane;
        const uint32_t o = i < nseg ? exec_owner_of(seg[i], world) : 0xffffffffu;
        for (uint32_t ob = 0; ob < world; ob += 32) {
            uint32_t add = 0;
            for (uint32_t k = 0; k < 32 && ob + k < world; k++) { uint32_t b = __ballot_sync(0xffffffffu, o == ob + k); if (lane == k) add = (uint32_t)__popc(b); }
            if (ob == 0) mine += add;
            else if (ob + lane < world && add) atomicAdd(&cnt[(uint64_t)(ob + lane) * nruns + run], add);   // rare: world > 32
        }
    }
    if (lane < world) cnt[(uint64_t)lane * nruns + run] = mine + (world > 32 ? cnt[(uint64_t)lane * nruns + run] : 0u);
}
// segment headers: count per owner from the scan (scan[o * nruns] = entries of all owners before o)
__global__ void k_xb_headers(const uint64_t* __restrict__ scan, const uint64_t* __restrict__ total, uint32_t world, uint32_t nruns, uint64_t cap, uint8_t* send,
                             unsigned long long* overflow) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= world) return;
    const uint64_t a = scan[(uint64_t)o * nruns], b = o + 1 < world ? scan[(uint64_t)(o + 1) * nruns] : *total;
    unsigned long long c = b - a;
    if (c > cap) { *overflow = 1; c = cap; }
    *(unsigned long long*)(send + (uint64_t)o * (XSEG_HDR + cap * 48)) = c;
}
// entries leave in (owner, position) order: slot = entries of that owner in earlier runs + earlier ones of this run
__global__ void __launch_bounds__(128) k_xb_scatter(const RawCid* __restrict__ seg, uint64_t nseg, uint64_t pos0, uint32_t world, uint32_t nruns,
                                                   const uint64_t* __restrict__ scan, uint64_t cap, uint8_t* send) {
    __shared__ uint32_t s_run[4][256];
    const uint32_t run = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (run >= nruns) return;
    for (uint32_t o = lane; o < world; o += 32) s_run[wib][o] = 0;
    __syncwarp();
    const uint64_t base = (uint64_t)run * XB_RUN;
    for (uint32_t c = 0; c < XB_RUN / 32; c++) {
        const uint64_t i = base + c * 32 + lane;
        const bool valid = i < nseg;
        RawCid rc{};
        uint32_t o = 0xffffffffu;
        if (valid) { rc = seg[i]; o = exec_owner_of(rc, world); }
        const unsigned m = __match_any_sync(0xffffffffu, o);
        if (valid) {
            const uint32_t rank = (uint32_t)__popc(m & ((1u << lane) - 1u));
            const uint64_t slot = scan[(uint64_t)o * nruns + run] - scan[(uint64_t)o * nruns] + s_run[wib][o] + rank;
            if (slot < cap) {
                ExecEntry e; e.c = rc; e.pos = pos0 + i;
                *(ExecEntry*)(send + (uint64_t)o * (XSEG_HDR + cap * 48) + XSEG_HDR + slot * 48) = e;
            }
        }
        __syncwarp();
        if (valid && (m & ((1u << lane) - 1u)) == 0) s_run[wib][o] += (uint32_t)__popc(m);   // first lane of every owner group
        __syncwarp();
    }
}

// seg_off[0..world] from the received segment headers
__global__ void k_recv_offsets(const uint8_t* __restrict__ recv, uint32_t world, uint64_t cap, uint64_t* seg_off) {
    if (threadIdx.x || blockIdx.x) return;
    uint64_t run = 0;
    for (uint32_t r = 0; r < world; r++) {
        seg_off[r] = run;
        uint64_t c = *(const unsigned long long*)(recv + (uint64_t)r * (XSEG_HDR + cap * 48));
        run += c > cap ? cap : c;
    }
    seg_off[world] = run;
}
__device__ __forceinline__ const ExecEntry* recv_entry_seg(const uint8_t* recv, const uint64_t* seg_off, uint32_t world, uint64_t cap, uint64_t k) {
    uint32_t r = 0;
    while (r + 1 < world && k >= seg_off[r + 1]) r++;
    return (const ExecEntry*)(recv + (uint64_t)r * (XSEG_HDR + cap * 48) + XSEG_HDR) + (k - seg_off[r]);
}
// one canonical slot per distinct CID holding the smallest entry ordinal (= smallest global position: segments arrive in rank
// order and are position-ordered inside)
__global__ void k_exec_claim_seg(const uint8_t* __restrict__ recv, const uint64_t* __restrict__ seg_off, uint32_t world, uint64_t cap,
                                 unsigned long long* table, uint64_t mask) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= seg_off[world]) return;
    const ExecEntry* e = recv_entry_seg(recv, seg_off, world, cap, k);
    uint64_t h = rawcid_hash(e->c);
    uint32_t fp = (uint32_t)(h >> 40) | 1u;
    unsigned long long mine = ((unsigned long long)fp << 32) | (unsigned long long)(k + 1);
    uint64_t slot = h & mask;
    for (;;) {
        unsigned long long v = table[slot];
        if (v == 0) { v = atomicCAS(&table[slot], 0ull, mine); if (v == 0) return; }
        if ((uint32_t)(v >> 32) == fp && rawcid_eq(recv_entry_seg(recv, seg_off, world, cap, (uint32_t)v - 1)->c, e->c)) { atomicMin(&table[slot], mine); return; }
        slot = (slot + 1) & mask;
    }
}
// every entry that is not the first occurrence of its CID sets the bit of its global position
__global__ void k_exec_mark_dups(const uint8_t* __restrict__ recv, const uint64_t* __restrict__ seg_off, uint32_t world, uint64_t cap,
                                 const unsigned long long* __restrict__ table, uint64_t mask, uint32_t* bitmap) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= seg_off[world]) return;
    const ExecEntry* e = recv_entry_seg(recv, seg_off, world, cap, k);
    uint64_t h = rawcid_hash(e->c);
    uint32_t fp = (uint32_t)(h >> 40) | 1u;
    uint64_t slot = h & mask;
    for (;;) {
        unsigned long long v = table[slot];
        if (v == 0) return;   // cannot happen: every entry was claimed
        if ((uint32_t)(v >> 32) == fp && rawcid_eq(recv_entry_seg(recv, seg_off, world, cap, (uint32_t)v - 1)->c, e->c)) {
            if ((uint32_t)v - 1 != (uint32_t)k) atomicOr(&bitmap[e->pos >> 5], 1u << (e->pos & 31));
            return;
        }
        slot = (slot + 1) & mask;
    }
}
// zero bits per word of the duplicate bitmap (positions past nraw do not count)
__global__ void k_zero_counts(const uint32_t* __restrict__ bitmap, uint64_t nraw, uint32_t* zeros) {
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t nwords = (nraw + 31) / 32;
    if (w >= nwords) return;
    uint32_t valid = (w == nwords - 1 && (nraw & 31)) ? ((1u << (nraw & 31)) - 1u) : 0xffffffffu;
    zeros[w] = (uint32_t)__popc(~bitmap[w] & valid);
}
// exec index i of every matching receipt → raw position of the (i+1)-th zero bit (UINT64_MAX past the end)
__global__ void k_select_positions(const uint32_t* __restrict__ match_rel, uint64_t n_match, uint64_t lo, const uint32_t* __restrict__ bitmap,
                                   const uint64_t* __restrict__ zprefix, uint64_t nwords, const unsigned long long* __restrict__ n_exec,
                                   uint64_t* out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_match) return;
    const uint64_t i = lo + match_rel[t];
    if (i >= *n_exec) { out[t] = ~0ull; return; }
    uint64_t a = 0, b = nwords;            // largest w with zprefix[w] <= i
    while (b - a > 1) { uint64_t m = (a + b) >> 1; if (zprefix[m] <= i) a = m; else b = m; }
    uint32_t x = ~bitmap[a];
    uint32_t k = (uint32_t)(i - zprefix[a]);
    for (uint32_t j = 0; j < k; j++) x &= x - 1;
    out[t] = a * 32 + (uint64_t)(__ffs((int)x) - 1);
}
__global__ void k_fetch_positions(const RawCid* __restrict__ seg, uint64_t nseg, uint64_t pos0, const uint64_t* __restrict__ req, uint64_t n, RawCid* out) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t p = req[j];
    RawCid z{};
    out[j] = (p >= pos0 && p - pos0 < nseg) ? seg[p - pos0] : z;
}
// EventProof.message_cid = exec[exec_index] (events/generator.rs:245, :289): answers are in the order of the matching list
__global__ void k_patch_message_cids(ipcfp_event_proof* proofs, uint64_t n_proofs, const uint32_t* __restrict__ match_rel, uint64_t n_match, uint64_t lo,
                                     const RawCid* __restrict__ answers) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_proofs) return;
    const uint64_t i = proofs[k].exec_index;
    if (i == 0xFFFFFFFFFFFFFFFFull || i < lo) return;
    const uint32_t rel = (uint32_t)(i - lo);
    uint64_t a = 0, b = n_match;
    while (b - a > 1) { uint64_t m = (a + b) >> 1; if (match_rel[m] <= rel) a = m; else b = m; }
    if (!n_match || match_rel[a] != rel) return;
    const RawCid c = answers[a];
    uint8_t* o = proofs[k].message_cid;
    for (int q = 0; q < 6; q++) o[q] = (uint8_t)(c.w[4] >> (8 * q));
    for (int q = 0; q < 32; q++) o[6 + q] = (uint8_t)(c.w[q >> 3] >> (8 * (q & 7)));
}
// n_exec = nraw − duplicates, published for pass 2 and the host
__global__ void k_set_n_exec(const uint64_t* __restrict__ zprefix_total, unsigned long long* n_exec) { *n_exec = *zprefix_total; }

// ------------------------------------------------------------------------------------------ witness union kernels
// 38-byte CIDs ↔ 40-byte records {digest[32], prefix[6], 0, 0} (aligned words for the merge)
__global__ void k_cids_to_recs(const uint8_t* __restrict__ cids, uint64_t n, uint64_t cap, RawCid* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cap) return;
    RawCid c{};
    if (i < n) {
        const uint8_t* s = cids + 38 * i;
        uint64_t pre = 0;
        for (int q = 0; q < 6; q++) pre |= (uint64_t)s[q] << (8 * q);
        c.w[4] = pre;
        for (int q = 0; q < 32; q++) c.w[q >> 3] |= (uint64_t)s[6 + q] << (8 * (q & 7));
    }
    out[i] = c;
}
__device__ __forceinline__ uint64_t bswap64_p(uint64_t x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}
// raw byte order of (prefix, digest) — `Cid` Ord for CIDs of one prefix (the Filecoin chain case, see ipcfp_merge_witness_cids)
__device__ __forceinline__ int rec_cmp(const RawCid& a, const RawCid& b) {
    uint64_t pa = bswap64_p(a.w[4] << 16), pb = bswap64_p(b.w[4] << 16);
    if (pa != pb) return pa < pb ? -1 : 1;
#pragma unroll
    for (int k = 0; k < 4; k++) { uint64_t x = bswap64_p(a.w[k]), y = bswap64_p(b.w[k]); if (x != y) return x < y ? -1 : 1; }
    return 0;
}
__device__ __forceinline__ uint32_t rec_bucket(const RawCid& a) { return (uint32_t)((a.w[0] & 0xff) << 8 | ((a.w[0] >> 8) & 0xff)); }
#define MERGE_BUCKETS 65536u
// starts[b][B - b0] = first index of list b whose bucket is >= B (B = b0..b0+nb); lists are sorted and hold buckets of [b0, b0+nb) only,
// so every element fills the gap it closes
__global__ void k_merge_starts(const RawCid* __restrict__ lists, const uint64_t* __restrict__ counts, uint32_t world, uint64_t cap, uint32_t* starts,
                               uint32_t b0, uint32_t nb) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)world * cap) return;
    uint32_t b = (uint32_t)(g / cap);
    uint64_t k = g % cap, n = counts[b];
    uint32_t* st = starts + (uint64_t)b * (nb + 1);
    if (n == 0) { if (k == 0) for (uint32_t B = 0; B <= nb; B++) st[B] = 0; return; }
    if (k >= n) return;
    const RawCid* L = lists + (uint64_t)b * cap;
    uint32_t Bk = rec_bucket(L[k]) - b0;
    uint32_t from = k == 0 ? 0 : rec_bucket(L[k - 1]) - b0 + 1;
    for (uint32_t B = from; B <= Bk; B++) st[B] = (uint32_t)k;
    if (k == n - 1) for (uint32_t B = Bk + 1; B <= nb; B++) st[B] = (uint32_t)n;
}
// position of every element in the merged (still non-unique) order + is it the first of its CID
__global__ void k_merge_rank(const RawCid* __restrict__ lists, const uint64_t* __restrict__ counts, uint32_t world, uint64_t cap,
                             const uint32_t* __restrict__ starts, uint32_t* pos_of, uint32_t* keep, uint32_t b0, uint32_t nb) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)world * cap) return;
    uint32_t b = (uint32_t)(g / cap);
    uint64_t k = g % cap;
    if (k >= counts[b]) return;
    const RawCid e = lists[(uint64_t)b * cap + k];
    const uint32_t B = rec_bucket(e) - b0;
    uint64_t pos = 0;
    bool dup = false;
    for (uint32_t q = 0; q < world; q++) {
        const uint32_t* st = starts + (uint64_t)q * (nb + 1);
        uint32_t s0 = st[B], s1 = st[B + 1];
        pos += s0;
        if (q == b) { pos += k - s0; continue; }
        const RawCid* L = lists + (uint64_t)q * cap;
        for (uint32_t x = s0; x < s1; x++) {
            int c = rec_cmp(L[x], e);
            if (c < 0 || (c == 0 && q < b)) pos++;
            if (c == 0 && q < b) dup = true;
            if (c > 0) break;
        }
    }
    pos_of[g] = (uint32_t)pos;
    keep[pos] = dup ? 0u : 1u;
}
__global__ void k_merge_emit38(const RawCid* __restrict__ lists, const uint64_t* __restrict__ counts, uint32_t world, uint64_t cap,
                               const uint32_t* __restrict__ pos_of, const uint32_t* __restrict__ keep, const uint64_t* __restrict__ outidx, uint8_t* out) {
    uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (uint64_t)world * cap) return;
    uint32_t b = (uint32_t)(g / cap);
    uint64_t k = g % cap;
    if (k >= counts[b]) return;
    uint32_t pos = pos_of[g];
    if (!keep[pos]) return;
    const RawCid c = lists[(uint64_t)b * cap + k];
    uint8_t* o = out + 38ull * outidx[pos];
    for (int q = 0; q < 6; q++) o[q] = (uint8_t)(c.w[4] >> (8 * q));
    for (int q = 0; q < 32; q++) o[6 + q] = (uint8_t)(c.w[q >> 3] >> (8 * (q & 7)));
}

// ---- partitioned union: rank r owns the CIDs whose bucket (first two digest bytes) lies in [part_lo(r), part_lo(r+1))
__host__ __device__ __forceinline__ uint32_t part_lo(uint32_t r, uint32_t world) { return (uint32_t)(((uint64_t)r * MERGE_BUCKETS + world - 1) / world); }
// One CTA: bounds[r] = first index of the sorted local list whose bucket is >= part_lo(r) (r = 0..world); the header record of piece r
// in the send buffer (piece stride = cap + 1 records) receives min(piece length, cap); *overflow = 1 when a piece does not fit.
__global__ void k_part_bounds(const RawCid* __restrict__ list, uint64_t n, uint32_t world, uint64_t cap, uint64_t* bounds, RawCid* send, unsigned long long* overflow) {
    __shared__ uint64_t b[1025];
    for (uint32_t r = threadIdx.x; r <= world; r += blockDim.x) {
        const uint32_t want = part_lo(r, world);
        uint64_t lo = 0, hi = n;
        while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (rec_bucket(list[mid]) < want) lo = mid + 1; else hi = mid; }
        b[r] = r == world ? n : lo;
        bounds[r] = b[r];
    }
    if (threadIdx.x == 0) *overflow = 0;
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < world; r += blockDim.x) {
        const uint64_t cnt = b[r + 1] - b[r];
        RawCid h{};
        h.w[0] = cnt < cap ? cnt : cap;
        send[(uint64_t)r * (cap + 1)] = h;
        if (cnt > cap) *overflow = 1;
    }
}
// entry i of the sorted local list → its place in the piece of the rank that owns its bucket
__global__ void k_part_pack(const RawCid* __restrict__ list, uint64_t n, uint32_t world, uint64_t cap, const uint64_t* __restrict__ bounds, RawCid* send) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RawCid e = list[i];
    const uint32_t dst = (uint32_t)(((uint64_t)rec_bucket(e) * world) >> 16);
    const uint64_t j = i - bounds[dst];
    if (j < cap) send[(uint64_t)dst * (cap + 1) + 1 + j] = e;
}
// piece lengths out of the received headers
__global__ void k_part_counts(const RawCid* __restrict__ recv, uint32_t world, uint64_t cap, uint64_t* counts) {
    for (uint32_t r = threadIdx.x; r < world; r += blockDim.x) counts[r] = recv[(uint64_t)r * (cap + 1)].w[0];
}
}  // namespace ipcfp

// ------------------------------------------------------------------------------------------ host side of the protocol
namespace ipcfp {

struct Words8 { uint64_t w[8]; };
__global__ void k_put_words(Words8 v, uint32_t k, unsigned long long* dst) { if (threadIdx.x < k) dst[threadIdx.x] = v.w[threadIdx.x]; }
__global__ void k_fill_words(unsigned long long* dst, uint32_t n, unsigned long long v) { for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = v; }
__global__ void k_publish_words(const unsigned long long* __restrict__ src, unsigned long long* dst_mapped, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dst_mapped[i] = src[i];
    __threadfence_system();
}
// word `field` of every rank's 8-word H2 record → dst[r]
__global__ void k_pick_field(const unsigned long long* __restrict__ gathered, uint32_t world, uint32_t stride, uint32_t field, uint64_t* dst) {
    for (uint32_t r = threadIdx.x; r < world; r += blockDim.x) dst[r] = gathered[(uint64_t)r * stride + field];
}
// Tiny all-gather of k ≤ 8 words per rank whose result the HOST needs. No copy-engine work (the D2H engine is busy with the
// witness blob): the words go up as a kernel parameter and come back through mapped host memory.
static void run_all_gather_host(Comm* c, const uint64_t* mine, uint32_t k, uint64_t* all /* = c->host.p */) {
    NcclApi* n = nccl_api();
    Words8 v{};
    for (uint32_t i = 0; i < k && i < 8; i++) v.w[i] = mine[i];
    k_put_words<<<1, 32, 0, c->sx>>>(v, k, c->words.p); IPCFP_LAUNCH_CHECK();
    IPCFP_NCCL(n->AllGather(c->words.p, c->words2.p, k, ncclUint64, c->cx, c->sx));
    k_publish_words<<<1, 256, 0, c->sx>>>(c->words2.p, (unsigned long long*)c->host.dev, c->world * k); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaStreamSynchronize(c->sx));
    (void)all;
}

uint64_t ShardExchange::host_word(uint32_t i) const { return s->host_words.p[i]; }
cudaStream_t ShardExchange::stream() const { return c->sx; }
cudaStream_t ShardExchange::union_stream() const { return c->sw; }
ShardExchange::ShardExchange(Comm* comm, Store* store, uint64_t lo_, uint64_t hi_) : c(comm), s(store), lo(lo_), hi(hi_) {
    if (c->device != s->device) throw Error(IPCFP_ERR_INVALID_ARG, "communicator and store are bound to different devices");
    if (c->world > 255) throw Error(IPCFP_ERR_UNSUPPORTED, "world too large");
}

// EARLY H0: can every shard promise the length of its slice already (dense message AMTs: known from the roots), and do all shards see the
// same message list. all_early ⇒ the slices (nseg_all, pos0, nraw) are set as for agree_slices.
void ShardExchange::agree_early(bool can_promise, uint64_t planned_nseg, uint64_t nraw_total) {
    const uint32_t W = c->world;
    uint64_t mine[4] = {can_promise ? 1ull : 0ull, planned_nseg, nraw_total, 0};
    uint64_t* all = c->host.p;
    run_all_gather_host(c, mine, 4, all);
    bool ok = true;
    nseg_all.assign(W, 0);
    nraw = 0; max_nseg = 0;
    for (uint32_t r = 0; r < W; r++) {
        if (!all[4 * r] || all[4 * r + 2] != all[2]) ok = false;
        nseg_all[r] = all[4 * r + 1];
        if (r == c->rank) pos0 = nraw;
        nraw += nseg_all[r];
        max_nseg = std::max(max_nseg, nseg_all[r]);
    }
    if (ok && nraw != all[2]) ok = false;                 // the promised slices must tile the whole list
    if (nraw >= 0xffffffffull) ok = false;
    all_early = ok;
    peers_ok = true;
    nseg = planned_nseg;
}
// H0: does every shard have its slice of the message list, and how long is it. Every rank takes part, also a failing one.
void ShardExchange::agree_slices(uint64_t tx_key, uint64_t err_key, uint64_t nseg_) {
    const uint32_t W = c->world;
    const bool ok = tx_key == IPCFP_NO_ERROR && err_key == IPCFP_NO_ERROR;
    uint64_t mine[4] = {ok ? 1ull : 0ull, ok ? nseg_ : 0, tx_key, err_key};
    uint64_t* all = c->host.p;
    run_all_gather_host(c, mine, 4, all);
    nseg_all.assign(W, 0);
    bool all_ok = true;
    nraw = 0; max_nseg = 0;
    g_tx = g_err = IPCFP_NO_ERROR;
    for (uint32_t r = 0; r < W; r++) {
        if (!all[4 * r]) all_ok = false;
        g_tx = std::min(g_tx, all[4 * r + 2]); g_err = std::min(g_err, all[4 * r + 3]);
        nseg_all[r] = all[4 * r + 1];
        if (r == c->rank) pos0 = nraw;
        nraw += nseg_all[r];
        max_nseg = std::max(max_nseg, nseg_all[r]);
    }
    peers_ok = all_ok;
    nseg = nseg_;
    if (nraw >= 0xffffffffull) throw Error(IPCFP_ERR_UNSUPPORTED, "more than 2^32 messages");
}

// X: bucketize → all-to-all → dedup → duplicate bitmap → all-reduce, all on the exchange stream (runs underneath pass 1)
void ShardExchange::start_exchange(const void* seg_dev, cudaEvent_t seg_ready) {
    NcclApi* n = nccl_api();
    const uint32_t W = c->world;
    cudaStream_t sx = c->sx;
    seg = (const RawCid*)seg_dev;
    IPCFP_CUDA(cudaStreamWaitEvent(sx, seg_ready, 0));
    IPCFP_CUDA(cudaEventRecord(c->tm[0], sx));
    cap = max_nseg / W + max_nseg / (4 * W) + 1024;
    const uint64_t segbytes = XSEG_HDR + cap * 48;
    c->sendbuf.ensure(segbytes * W);
    c->recvbuf.ensure(segbytes * W);
    nwords = (nraw + 31) / 32;
    c->bitmap.ensure(nwords + 64);
    c->bitmap_sum.ensure(nwords + 64);
    c->zeros.ensure(nwords + 64);
    c->zprefix.ensure(nwords + 64);
    c->scan_tmp.ensure(scan_scratch_elems(nwords + 64) + 64);
    uint64_t slots = 64;
    while (slots < 2 * (W * cap)) slots <<= 1;
    c->table.ensure(slots);
    unsigned long long* overflow = c->words.p + 3100;
    IPCFP_CUDA(cudaMemsetAsync(c->bitmap.p, 0, (nwords + 64) * 4, sx));
    IPCFP_CUDA(cudaMemsetAsync(c->table.p, 0, slots * 8, sx));
    IPCFP_CUDA(cudaMemsetAsync(overflow, 0, 8, sx));
    // headers of empty segments must read 0 even when this rank has nothing to send
    for (uint32_t r = 0; r < W; r++) IPCFP_CUDA(cudaMemsetAsync(c->sendbuf.p + r * segbytes, 0, XSEG_HDR, sx));
    if (nseg) {
        const uint32_t nruns = div_up(nseg, XB_RUN);
        AsyncBuf<uint32_t> cnt((uint64_t)W * nruns + 64, sx);
        AsyncBuf<uint64_t> scan((uint64_t)W * nruns + 64, sx), scratch(scan_scratch_elems((uint64_t)W * nruns) + 8, sx), total(1, sx);
        if (W > 32) cnt.zero();
        k_xb_count<<<div_up((uint64_t)nruns * 32, 128), 128, 0, sx>>>(seg, nseg, W, nruns, cnt.p); IPCFP_LAUNCH_CHECK();
        exclusive_scan_u32(cnt.p, scan.p, (uint64_t)W * nruns, total.p, scratch.p, sx);
        k_xb_headers<<<div_up(W, 64), 64, 0, sx>>>(scan.p, total.p, W, nruns, cap, c->sendbuf.p, overflow); IPCFP_LAUNCH_CHECK();
        k_xb_scatter<<<div_up((uint64_t)nruns * 32, 128), 128, 0, sx>>>(seg, nseg, pos0, W, nruns, scan.p, cap, c->sendbuf.p); IPCFP_LAUNCH_CHECK();
    }
    IPCFP_CUDA(cudaEventRecord(c->tm[6], sx));
    // all-to-all of whole segments (fixed size: no count round trip; the valid count travels in the segment header)
    IPCFP_NCCL(n->GroupStart());
    for (uint32_t p = 0; p < W; p++) {
        IPCFP_NCCL(n->Send(c->sendbuf.p + p * segbytes, segbytes, ncclUint8, (int)p, c->cx, sx));
        IPCFP_NCCL(n->Recv(c->recvbuf.p + p * segbytes, segbytes, ncclUint8, (int)p, c->cx, sx));
    }
    IPCFP_NCCL(n->GroupEnd());
    IPCFP_CUDA(cudaEventRecord(c->tm[7], sx));
    uint64_t* seg_off = (uint64_t*)(c->words.p + 2600);   // [0, W]
    k_recv_offsets<<<1, 1, 0, sx>>>(c->recvbuf.p, W, cap, seg_off); IPCFP_LAUNCH_CHECK();
    const unsigned g = div_up(W * cap, 256);
    k_exec_claim_seg<<<g, 256, 0, sx>>>(c->recvbuf.p, seg_off, W, cap, c->table.p, slots - 1); IPCFP_LAUNCH_CHECK();
    k_exec_mark_dups<<<g, 256, 0, sx>>>(c->recvbuf.p, seg_off, W, cap, c->table.p, slots - 1, c->bitmap.p); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaEventRecord(c->tm[8], sx));
    // the owners' bitmaps are disjoint (a position belongs to one CID, a CID to one owner): their sum is their union
    IPCFP_NCCL(n->AllReduce(c->bitmap.p, c->bitmap_sum.p, nwords + 1, ncclUint32, ncclSum, c->cx, sx));
    // n_exec = number of zero bits; prefix zero counts for the select
    if (nwords) { k_zero_counts<<<div_up(nwords, 256), 256, 0, sx>>>(c->bitmap_sum.p, nraw, c->zeros.p); IPCFP_LAUNCH_CHECK(); }
    n_exec_dev = c->words.p + 3101;
    exclusive_scan_u32(c->zeros.p, c->zprefix.p, nwords, (uint64_t*)n_exec_dev, c->scan_tmp.p, sx);
    IPCFP_CUDA(cudaEventRecord(c->ev_a, sx));
    IPCFP_CUDA(cudaEventRecord(c->tm[1], sx));
    overflow_dev = overflow;
}

// P: before pass 2 — the global n_exec for the "Missing message at index" check, raw positions of this rank's matches
void ShardExchange::positions_for(cudaStream_t st, const uint32_t* match_rel, uint64_t n_match, unsigned long long* n_exec_out) {
    IPCFP_CUDA(cudaStreamWaitEvent(st, c->ev_a, 0));
    IPCFP_CUDA(cudaMemcpyAsync(n_exec_out, n_exec_dev, 8, cudaMemcpyDeviceToDevice, st));
    publish_words_on(s, st, overflow_dev, 300, 2);   // host_words[300] = exchange overflow flag, [301] = n_exec: read after the next sync of that stream
    M = n_match;
    c->req.ensure(n_match + 64);
    if (n_match) {
        k_select_positions<<<div_up(n_match, 128), 128, 0, st>>>(match_rel, n_match, lo, c->bitmap_sum.p, c->zprefix.p, nwords, n_exec_dev, c->req.p);
        IPCFP_LAUNCH_CHECK();
    }
    match_rel_dev = match_rel;
}

// H2: every rank reports how far it got; all ranks continue or fail TOGETHER, with the same (first) error
void ShardExchange::agree_results(uint64_t tx_key, uint64_t err_key, bool missing_base, uint64_t n_proofs, uint64_t n_witness, uint64_t exch_overflow, bool stale) {
    const uint32_t W = c->world;
    uint64_t mine[8] = {tx_key, err_key, missing_base ? 1ull : 0ull, M, n_proofs, n_witness, exch_overflow, stale ? 1ull : 0ull};
    uint64_t* all = c->host.p;
    run_all_gather_host(c, mine, 8, all);
    k_pick_field<<<1, 256, 0, c->sx>>>(c->words2.p, W, 8, 5, (uint64_t*)(c->words2.p + 2600)); IPCFP_LAUNCH_CHECK();
    IPCFP_CUDA(cudaStreamSynchronize(c->sx));
    g_tx = g_err = IPCFP_NO_ERROR; g_missing_base = false; g_overflow = false; g_stale = false;
    M_max = 0; nw_max = 0; M_total = 0; proofs_total = 0;
    nw_all.assign(W, 0);
    for (uint32_t r = 0; r < W; r++) {
        const uint64_t* a = all + 8 * r;
        g_tx = std::min(g_tx, a[0]); g_err = std::min(g_err, a[1]);
        g_missing_base |= a[2] != 0;
        M_max = std::max(M_max, a[3]); M_total += a[3]; proofs_total += a[4];
        nw_all[r] = a[5]; nw_max = std::max(nw_max, a[5]);
        g_overflow |= a[6] != 0;
        g_stale |= a[7] != 0;
    }
}

// F: positions wanted by every rank → the owners answer → EventProof.message_cid of this rank's proofs
void ShardExchange::fetch_and_patch(cudaStream_t st, ipcfp_event_proof* proofs_dev, uint64_t n_proofs) {
    IPCFP_CUDA(cudaEventRecord(c->tm[2], st));
    IPCFP_CUDA(cudaEventRecord(c->tm[3], st));
    if (M_max == 0) return;
    NcclApi* n = nccl_api();
    const uint32_t W = c->world;
    c->req_all.ensure((uint64_t)W * M_max + 64);
    c->ans.ensure(((uint64_t)W * M_max + 8) * 5);
    c->ans_sum.ensure(((uint64_t)W * M_max + 8) * 5);
    // this rank's request list, padded to M_max with "nobody's position" (req itself holds M entries: M_max was not known when it was sized)
    c->req_pad.ensure(M_max + 64);
    IPCFP_CUDA(cudaMemsetAsync(c->req_pad.p, 0xff, M_max * 8, st));
    if (M) IPCFP_CUDA(cudaMemcpyAsync(c->req_pad.p, c->req.p, M * 8, cudaMemcpyDeviceToDevice, st));
    IPCFP_NCCL(n->AllGather(c->req_pad.p, c->req_all.p, M_max, ncclUint64, c->cx, st));
    const uint64_t total = (uint64_t)W * M_max;
    k_fetch_positions<<<div_up(total, 256), 256, 0, st>>>(seg, nseg, pos0, c->req_all.p, total, (RawCid*)c->ans.p); IPCFP_LAUNCH_CHECK();
    IPCFP_NCCL(n->AllReduce(c->ans.p, c->ans_sum.p, total * 5, ncclUint64, ncclSum, c->cx, st));
    if (n_proofs) {
        k_patch_message_cids<<<div_up(n_proofs, 128), 128, 0, st>>>(proofs_dev, n_proofs, match_rel_dev, M, lo, (const RawCid*)c->ans_sum.p + (uint64_t)c->rank * M_max);
        IPCFP_LAUNCH_CHECK();
    }
    IPCFP_CUDA(cudaEventRecord(c->tm[3], st));
}

// W: union of the per-shard sorted witness CID lists (BTreeSet union of common/witness.rs:24-40) on every rank
void ShardExchange::witness_union(cudaStream_t st, const uint8_t* cids_dev, uint64_t n_local, uint8_t** out_dev, uint64_t* n_out_dev_word) {
    NcclApi* n = nccl_api();
    const uint32_t W = c->world;
    IPCFP_CUDA(cudaEventRecord(c->tm[4], st));
    const uint64_t capw = nw_max + 1;
    const uint64_t total_cap = (uint64_t)W * capw;
    c->recs.ensure(capw * 40 + 64);
    c->gather.ensure(total_cap * 40 + 64);
    c->starts.ensure((uint64_t)W * (MERGE_BUCKETS + 1) + 64);
    c->pos_of.ensure(total_cap + 64);
    c->flags.ensure(total_cap + 64);
    c->fscan.ensure(total_cap + 64);
    c->merged.ensure(total_cap * 38 + 64);
    c->scan_tmp.ensure(scan_scratch_elems(total_cap + 64) + 64);
    k_cids_to_recs<<<div_up(capw, 256), 256, 0, st>>>(cids_dev, n_local, capw, (RawCid*)c->recs.p); IPCFP_LAUNCH_CHECK();
    IPCFP_NCCL(n->AllGather(c->recs.p, c->gather.p, capw * 40, ncclUint8, c->cw, st));
    uint64_t* counts = (uint64_t*)(c->words2.p + 2600);   // per-rank list lengths, picked out of the H2 records by agree_results
    IPCFP_CUDA(cudaMemsetAsync(c->flags.p, 0, (total_cap + 64) * 4, st));
    const unsigned g = div_up(total_cap, 256);
    k_merge_starts<<<g, 256, 0, st>>>((const RawCid*)c->gather.p, counts, W, capw, c->starts.p, 0, MERGE_BUCKETS); IPCFP_LAUNCH_CHECK();
    k_merge_rank<<<g, 256, 0, st>>>((const RawCid*)c->gather.p, counts, W, capw, c->starts.p, c->pos_of.p, c->flags.p, 0, MERGE_BUCKETS); IPCFP_LAUNCH_CHECK();
    uint64_t total_listed = 0;
    for (uint32_t r = 0; r < W; r++) total_listed += nw_all[r];
    unsigned long long* n_union = c->words2.p + 3000;
    exclusive_scan_u32(c->flags.p, c->fscan.p, total_listed, (uint64_t*)n_union, c->scan_tmp.p, st);
    k_merge_emit38<<<g, 256, 0, st>>>((const RawCid*)c->gather.p, counts, W, capw, c->pos_of.p, c->flags.p, c->fscan.p, c->merged.p); IPCFP_LAUNCH_CHECK();
    *out_dev = c->merged.p;
    IPCFP_CUDA(cudaMemcpyAsync(n_out_dev_word, n_union, 8, cudaMemcpyDeviceToDevice, st));
    IPCFP_CUDA(cudaEventRecord(c->tm[5], st));
}
// W, partitioned: the union is left DISTRIBUTED — rank r ends up with the sorted, duplicate-free CIDs whose first two digest bytes
// fall into its 1/world share of the 65 536 buckets, so the concatenation of the partitions in rank order is the BTreeSet order. Every
// rank sends each peer the piece of its sorted list that belongs to it and merges the `world` sorted pieces it receives: bytes on the
// wire and merge work per rank are those of about TWO shard lists, whatever the world size (the all-gather variant above moves and
// merges `world` lists on every rank). Pieces travel in fixed-size slots of `cap` records behind a count header, so that no size has
// to come back to the host in the middle of the protocol; a piece that does not fit raises this rank's overflow word, every rank
// sees every word after the call's last synchronisation, and the caller repeats the union with cap = the longest list (cannot
// overflow). No host synchronisation inside.
uint64_t ShardExchange::union_piece_cap(bool cannot_overflow) const {
    const uint32_t W = c->world;
    if (cannot_overflow) return nw_max + 1;
    if (const char* e = getenv("IPCFP_UNION_CAP")) return (uint64_t)std::max(1, atoi(e));   // tests: force the overflow path
    return std::min<uint64_t>(nw_max + 1, 2 * ((nw_max + W - 1) / W) + 1024);
}
void ShardExchange::witness_union_partitioned(cudaStream_t st, const uint8_t* cids_dev, uint64_t n_local, uint64_t cap, uint8_t** out_dev, uint32_t host_word_first) {
    NcclApi* n = nccl_api();
    const uint32_t W = c->world, me = c->rank;
    IPCFP_CUDA(cudaEventRecord(c->tm[4], st));
    const uint64_t stride = cap + 1, total_cap = (uint64_t)W * stride;
    const uint32_t b0 = part_lo(me, W), nb = part_lo(me + 1, W) - b0;
    c->recs.ensure((n_local + 1) * 40 + 64);
    c->part_send.ensure(total_cap * 40 + 64);
    c->gather.ensure(total_cap * 40 + 64);
    c->part_words.ensure(4ull * W + 16);
    c->starts.ensure((uint64_t)W * (nb + 1) + 64);
    c->pos_of.ensure(total_cap + 64);
    c->flags.ensure(total_cap + 64);
    c->fscan.ensure(total_cap + 64);
    c->merged.ensure(total_cap * 38 + 64);
    c->scan_tmp.ensure(scan_scratch_elems(total_cap + 64) + 64);
    unsigned long long* pw = c->part_words.p;
    unsigned long long *bounds_d = pw, *col_d = pw + W + 1, *mine_d = pw + 2 * W + 2 /* [n_part, overflow] */, *all_d = pw + 2 * W + 4 /* 2 words per rank */;
    RawCid* send = (RawCid*)c->part_send.p;
    RawCid* recv = (RawCid*)c->gather.p;
    k_cids_to_recs<<<div_up(n_local + 1, 256), 256, 0, st>>>(cids_dev, n_local, n_local + 1, (RawCid*)c->recs.p); IPCFP_LAUNCH_CHECK();
    k_part_bounds<<<1, 256, 0, st>>>((const RawCid*)c->recs.p, n_local, W, cap, (uint64_t*)bounds_d, send, mine_d + 1); IPCFP_LAUNCH_CHECK();
    if (n_local) { k_part_pack<<<div_up(n_local, 256), 256, 0, st>>>((const RawCid*)c->recs.p, n_local, W, cap, (const uint64_t*)bounds_d, send); IPCFP_LAUNCH_CHECK(); }
    IPCFP_NCCL(n->GroupStart());
    for (uint32_t r = 0; r < W; r++) {
        IPCFP_NCCL(n->Send(send + (uint64_t)r * stride, stride * 40, ncclUint8, (int)r, c->cw, st));
        IPCFP_NCCL(n->Recv(recv + (uint64_t)r * stride, stride * 40, ncclUint8, (int)r, c->cw, st));
    }
    IPCFP_NCCL(n->GroupEnd());
    IPCFP_CUDA(cudaMemsetAsync(c->flags.p, 0, (total_cap + 64) * 4, st));
    k_part_counts<<<1, 256, 0, st>>>(recv, W, cap, (uint64_t*)col_d); IPCFP_LAUNCH_CHECK();
    const unsigned g = div_up(total_cap, 256);
    const RawCid* lists = recv + 1;   // piece r's entries start one record behind its header: same stride
    k_merge_starts<<<g, 256, 0, st>>>(lists, (const uint64_t*)col_d, W, stride, c->starts.p, b0, nb); IPCFP_LAUNCH_CHECK();
    k_merge_rank<<<g, 256, 0, st>>>(lists, (const uint64_t*)col_d, W, stride, c->starts.p, c->pos_of.p, c->flags.p, b0, nb); IPCFP_LAUNCH_CHECK();
    exclusive_scan_u32(c->flags.p, c->fscan.p, total_cap, (uint64_t*)mine_d, c->scan_tmp.p, st);
    k_merge_emit38<<<g, 256, 0, st>>>(lists, (const uint64_t*)col_d, W, stride, c->pos_of.p, c->flags.p, c->fscan.p, c->merged.p); IPCFP_LAUNCH_CHECK();
    // [partition size, overflow] of all ranks → the store's mapped words [host_word_first, +2·world): read after the caller's next sync of this stream
    IPCFP_NCCL(n->AllGather(mine_d, all_d, 2, ncclUint64, c->cw, st));
    publish_words_on(s, st, all_d, host_word_first, 2 * W);
    *out_dev = c->merged.p;
    IPCFP_CUDA(cudaEventRecord(c->tm[5], st));
}
void ShardExchange::timings(float* ms_exchange, float* ms_fetch, float* ms_union) const {
    cudaEventElapsedTime(ms_exchange, c->tm[0], c->tm[1]);
    if (getenv("IPCFP_XCH_TRACE")) {
        float a = 0, b = 0, d = 0, e = 0;
        cudaEventElapsedTime(&a, c->tm[0], c->tm[6]); cudaEventElapsedTime(&b, c->tm[6], c->tm[7]); cudaEventElapsedTime(&d, c->tm[7], c->tm[8]); cudaEventElapsedTime(&e, c->tm[8], c->tm[1]);
        fprintf(stderr, "[ipcfp rank %u] exchange: bucketize %.3f ms | all-to-all %.3f ms | dedup %.3f ms | all-reduce + scan %.3f ms (cap %llu entries/segment)\n", c->rank, a, b, d, e,
                (unsigned long long)cap);
    }
    cudaEventElapsedTime(ms_fetch, c->tm[2], c->tm[3]);
    cudaEventElapsedTime(ms_union, c->tm[4], c->tm[5]);
}
// IPCFP_XCH_TRACE: where the protocol's stages sit on the call's own time axis (ms after `origin`, an event of the engine stream)
void ShardExchange::trace_timeline(cudaEvent_t origin, const char* engine_part) const {
    float t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) cudaEventElapsedTime(&t[i], origin, c->tm[i]);
    fprintf(stderr, "[ipcfp rank %u] timeline ms: %s | exchange %.3f-%.3f fetch %.3f-%.3f union %.3f-%.3f\n", c->rank, engine_part, t[0], t[1], t[2], t[3], t[4], t[5]);
}

}  // namespace ipcfp
