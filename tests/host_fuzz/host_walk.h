// host_walk.h — the GENERAL message-AMT walk (count → scan → expand per level; any AMT shape, exact errors) driven on the host with
// the product's per-item functions `amt_item_count` / `amt_item_expand` (csrc/walk.cuh); mirrors the launch loop `run_general` of
// generate_event_proof (csrc/events.cu). TEST INFRASTRUCTURE ONLY. Include after host_shims.h + csrc/walk.cuh.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace ipcfp {

// roots: frontier seeds as k_setup writes them. rlo / rhi: per-AMT index ranges (shard_amt_ranges). Errors go to *err (the message-AMT fault word, tx_err_key).
// On return vals holds the raw execution list of the share (nraw entries).
inline void host_general_walk(const StoreView& sv, uint32_t namt, const std::vector<uint32_t>& f_blk, const std::vector<uint32_t>& f_meta, uint32_t last_round,
                              const uint64_t* rlo, const uint64_t* rhi, uint32_t record, uint32_t* wbits, unsigned long long* err, uint64_t cap,
                              std::vector<RawCid>& vals, uint64_t& nraw) {
    std::vector<uint32_t> A_blk(cap), A_meta(cap), B_blk(cap), B_meta(cap), counts(cap + 1);
    std::vector<uint64_t> A_base(cap, 0), B_base(cap, 0), out_off(cap + 1);
    std::copy(f_blk.begin(), f_blk.end(), A_blk.begin());
    std::copy(f_meta.begin(), f_meta.end(), A_meta.begin());
    Frontier fcur{A_blk.data(), A_meta.data(), A_base.data()}, fnxt{B_blk.data(), B_meta.data(), B_base.data()};
    unsigned long long cnt = namt;
    for (uint32_t round = 0; round <= last_round; round++) {
        const uint64_t items = std::min<uint64_t>(cnt, cap);
        uint64_t total = 0;
        for (uint64_t t = 0; t < items; t++) {
            counts[t] = amt_item_count(sv, fcur.blk[t], fcur.meta[t], fcur.base[t], round, last_round, rlo, rhi);
            out_off[t] = total;
            total += counts[t];
        }
        if (round == last_round) vals.assign(total + 8, RawCid{});
        ExpandArgs a;
        memset(&a, 0, sizeof a);
        a.store = sv; a.in = fcur; a.in_count = &cnt; a.out_off = out_off.data(); a.round = round; a.last_round = last_round; a.record = record;
        a.wbits = wbits; a.err = err; a.out = fnxt; a.vals = vals.data(); a.cap = (uint32_t)cap; a.rlo = rlo; a.rhi = rhi;
        for (uint64_t t = 0; t < items; t++)
            for (uint32_t j = 0; j < 8; j++) amt_item_expand(a, t, j, fcur.blk[t], fcur.meta[t], fcur.base[t], counts[t]);
        unsigned long long n = total;
        if (round < last_round && n > cap) { report_tx_error(err, IPCFP_TX_EIDX_NONE, 0, 0, DC_UNSUPPORTED, 1); n = cap; }
        cnt = n;
        std::swap(fcur, fnxt);
    }
    nraw = cnt;
}

}  // namespace ipcfp
