// host_store.h — a host copy of the engine's block store, laid out exactly as ipcfp_store_create does (csrc/store.cu): arena with
// lead / tail padding, one BlockRec per block, open-addressing CID index (equal CIDs keep the smallest index), one CID class.
// TEST INFRASTRUCTURE ONLY. Include AFTER csrc/store.cuh (or any header that includes it).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace ipcfp {

struct HostStore {
    std::vector<uint8_t> arena;
    std::vector<BlockRec> recs;
    std::vector<uint64_t> table;
    StoreView view;
    // mirrors ipcfp_store_create (csrc/store.cu): one CID class, digests from the CID bytes, equal CIDs keep the smallest index
    HostStore(const uint8_t* cids, const uint64_t* offs, const uint32_t* lens, const uint8_t* blob, uint64_t blob_size, uint64_t n) {
        arena.assign(16 + blob_size + 32, 0);
        memcpy(arena.data() + 16, blob, blob_size);
        recs.resize(n);
        memset(&view, 0, sizeof view);
        view.n_classes = 1;
        memcpy(view.class_prefix[0], cids, 6);
        uint64_t slots = 64;
        while (slots < 2 * n) slots <<= 1;
        table.assign(slots, 0);
        for (uint64_t i = 0; i < n; i++) {
            if (memcmp(cids + 38 * i, cids, 6)) { fprintf(stderr, "emu: several CID classes are not modelled\n"); exit(2); }
            BlockRec r;
            memset(&r, 0, sizeof r);
            memcpy(r.d.w, cids + 38 * i + 6, 32);
            r.off = offs[i]; r.len = lens[i]; r.cls = 0;
            recs[i] = r;
            uint64_t h = digest_hash(r.d, 0);
            uint32_t fp = (uint32_t)(h >> 32) | 1u;
            uint64_t slot = h & (slots - 1);
            for (;;) {
                uint64_t e = table[slot];
                if (e == 0) { table[slot] = ((uint64_t)fp << 32) | (i + 1); break; }
                if ((uint32_t)(e >> 32) == fp && digest_eq(recs[(uint32_t)e - 1].d, r.d)) break;   // first occurrence stays
                slot = (slot + 1) & (slots - 1);
            }
        }
        view.blob = arena.data() + 16;
        view.recs = recs.data();
        view.table = table.data();
        view.mask = slots - 1;
        view.n = (uint32_t)n;
    }
};


// store_lookup on a CID that lives in an exactly-38-byte HOST buffer. The device code loads CIDs with aligned 8-byte words and funnel
// shifts (load_digest, common.cuh), i.e. it may touch up to 7 bytes either side of the 38; every DEVICE buffer a kernel reads CIDs from
// is padded for that (arena pads, `+ 64` on the tipset arrays). The emulation gives the same guarantee with a padded copy, so that the
// harnesses can run under AddressSanitizer and everything it still reports is a real out-of-bounds access of the device code.
static inline int32_t store_lookup_host_cid(const StoreView& sv, const uint8_t* cid38) {
    alignas(16) uint8_t pad[64] = {0};
    memcpy(pad + 8, cid38, 38);
    return store_lookup(sv, pad + 8);
}
}  // namespace ipcfp
