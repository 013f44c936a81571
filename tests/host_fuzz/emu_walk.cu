// emu_walk.cu — the dense message-AMT walk executed ON THE CPU (TEST INFRASTRUCTURE, no GPU needed).
//
// csrc/walk.cuh holds the per-item device function of the walk (`amt_item_dense`) and the host-side planning
// (`shard_amt_ranges`, `make_dense_plan`). This program compiles them for the host, builds synthetic tipsets, lays the block
// store out exactly as ipcfp_store_create does (arena with pads, BlockRec array, open-addressing CID index), runs the walk
// level by level, item by item, lane by lane — for the whole tipset and for every shard of several world sizes — and
// compares with the oracle: the raw execution list (slice of the oracle's concatenated message list) and the recorded blocks
// (the oracle's shard witness for a spec that matches nothing = base witness + the shard's message-AMT blocks).
//
//   nvcc -std=c++17 -O2 -o emu_walk tests/host_fuzz/emu_walk.cu oracle/oracle.cpp synth/synth.cpp -lpthread && ./emu_walk
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "host_shims.h"

#include "../../ipc_filecoin_proofs_b200/csrc/walk.cuh"
#include "../../oracle/oracle.h"
#include "../../synth/synth.h"
#include "host_store.h"
#include "host_walk.h"

using namespace ipcfp;

static int fail(const char* what, uint64_t a = 0, uint64_t b = 0) {
    fprintf(stderr, "EMU MISMATCH: %s (%llu, %llu)\n", what, (unsigned long long)a, (unsigned long long)b);
    return 1;
}

static int run_case(const synth_params& sp, uint32_t world, uint64_t* walked_nodes) {
    synth_tipset* ts = synth_build(&sp);
    const uint64_t n = synth_n_blocks(ts);
    const uint8_t* cids = synth_cids(ts);
    HostStore hs(cids, synth_offsets(ts), synth_lengths(ts), synth_blob(ts), synth_blob_size(ts), n);
    const StoreView& sv = hs.view;
    oracle_store* os = oracle_store_create(cids, synth_offsets(ts), synth_lengths(ts), synth_blob(ts), n);
    ipcfp_tipset_desc td;
    memset(&td, 0, sizeof td);
    td.parent_epoch = synth_parent_epoch(ts); td.child_epoch = synth_child_epoch(ts); td.n_parents = synth_n_parents(ts);
    td.parent_cids = synth_parent_cids(ts); td.parent_txmeta_cids = synth_parent_txmeta_cids(ts); td.child_cid = synth_child_cid(ts);
    td.receipts_root = synth_receipts_root(ts); td.child_parent_state_root = synth_parent_state_root(ts); td.n_receipts = synth_n_receipts(ts);
    td.events_roots = synth_events_roots(ts); td.has_events_root = synth_has_events_root(ts);
    const uint32_t P = td.n_parents, namt = 2 * P;
    // the oracle's raw list
    std::vector<uint8_t> raw38(38ull * (td.n_receipts + 64ull * P + 1024));
    uint64_t nraw_oracle = 0;
    if (oracle_message_list(os, &td, raw38.data(), raw38.size() / 38, &nraw_oracle) != IPCFP_OK) return fail("oracle_message_list failed");
    int rc = 0;
    for (uint32_t rank = 0; rank < world && !rc; rank++) {
        const bool sharded = world > 1;
        const uint64_t lo = td.n_receipts * rank / world, hi = td.n_receipts * (rank + 1) / world;
        // ---- what k_setup does (csrc/events.cu): base witness marks, TxMeta → AMT roots → frontier seeds
        std::vector<uint32_t> wbits((n + 31) / 32 + 8, 0);
        auto mark_cid = [&](const uint8_t* cid) { int32_t b = store_lookup_host_cid(sv, cid); if (b < 0) return false; witness_mark(wbits.data(), (uint32_t)b); return true; };
        for (uint32_t b = 0; b < P; b++) if (!mark_cid(td.parent_cids + 38 * b) || !mark_cid(td.parent_txmeta_cids + 38 * b)) return fail("base block missing");
        if (!mark_cid(td.child_cid) || !mark_cid(td.receipts_root)) return fail("base block missing");
        std::vector<uint32_t> heights(namt), f_blk(namt), f_meta(namt);
        std::vector<uint64_t> counts(namt), f_base(namt, 0);
        for (uint32_t b = 0; b < P; b++) {
            int32_t tb = store_lookup_host_cid(sv, td.parent_txmeta_cids + 38 * b);
            uint32_t len;
            const uint8_t* p = store_block(sv, (uint32_t)tb, len);
            Rd r(p, len);
            rd_array_exact(r, 2);
            uint32_t c0 = rd_cid(r), c1 = rd_cid(r);
            rd_end(r);
            if (r.err) return fail("TxMeta decode");
            for (uint32_t k = 0; k < 2; k++) {
                int32_t rb = store_lookup(sv, p + (k ? c1 : c0));
                if (rb < 0) return fail("AMT root missing");
                witness_mark(wbits.data(), (uint32_t)rb);
                uint32_t rl;
                const uint8_t* rp = store_block(sv, (uint32_t)rb, rl);
                Rd rr(rp, rl);
                uint32_t bw, h;
                uint64_t cnt;
                amt_root_begin(rr, 0, bw, h, cnt);
                if (rr.err) return fail("AMT root decode");
                const uint32_t amt = 2 * b + k;
                f_blk[amt] = (uint32_t)rb; f_meta[amt] = make_meta(amt, 1, h); heights[amt] = h; counts[amt] = cnt;
            }
        }
        // ---- host planning (the real functions)
        std::vector<uint64_t> rlo(namt), rhi(namt);
        const uint64_t nraw_total = shard_amt_ranges(namt, counts.data(), sharded, lo, hi, td.n_receipts, rlo.data(), rhi.data());
        if (nraw_total != nraw_oracle) return fail("Nraw differs from the oracle", nraw_total, nraw_oracle);
        DensePlan plan = make_dense_plan(namt, heights.data(), counts.data(), rlo.data(), rhi.data(), 4 * n + 1024, 8 * (4 * n + 1024), 32768);   // the engine's own limits
        if (!plan.ok) continue;                               // geometry left to the general walk (e.g. a shard without messages)
        // ---- the walk: rounds × items × 8 lanes, exactly the kernel's indexing
        uint64_t fmax = 1;
        for (uint32_t r = 0; r < plan.rounds; r++) fmax = std::max<uint64_t>(fmax, plan.ftot[r]);
        std::vector<uint32_t> A_blk(fmax), A_meta(fmax), B_blk(fmax), B_meta(fmax), flen(2 * fmax + 8);
        std::vector<uint64_t> A_base(fmax), B_base(fmax), foff(2 * fmax + 8);
        std::copy(f_blk.begin(), f_blk.end(), A_blk.begin());
        std::copy(f_meta.begin(), f_meta.end(), A_meta.begin());
        std::copy(f_base.begin(), f_base.end(), A_base.begin());
        std::vector<RawCid> vals(plan.nraw + 8);
        uint32_t failflag = 0;
        DenseArgs a;
        memset(&a, 0, sizeof a);
        a.store = sv;
        a.ping = Frontier{A_blk.data(), A_meta.data(), A_base.data()};
        a.pong = Frontier{B_blk.data(), B_meta.data(), B_base.data()};
        a.vals = vals.data();
        a.vbase = plan.per_amt.data(); a.cnt = plan.per_amt.data() + namt; a.lo = plan.per_amt.data() + 2ull * namt; a.hi = plan.per_amt.data() + 3ull * namt;
        a.fofs = plan.fofs.data(); a.ftot = plan.ftot.data();
        a.namt = namt; a.record = 1; a.wbits = wbits.data(); a.fail = &failflag;
        a.f_off[0] = foff.data(); a.f_off[1] = foff.data() + fmax; a.f_len[0] = flen.data(); a.f_len[1] = flen.data() + fmax;
        for (uint32_t round = 0; round < plan.rounds && !failflag; round++) {
            const Frontier in = (round & 1) ? a.pong : a.ping, out = (round & 1) ? a.ping : a.pong;
            for (uint32_t it = 0; it < plan.ftot[round]; it++)
                for (uint32_t j = 0; j < 8; j++) amt_item_dense(a, in, out, round, it, j);
            *walked_nodes += plan.ftot[round];
        }
        if (failflag) return fail("the dense walk raised its flag on a well-formed dense tipset", rank, world);
        // ---- raw list == the oracle's slice
        // (a tipset without receipts has no shares to own: shard_amt_ranges, csrc/walk.cuh, gives every rank the empty range)
        const uint64_t glo = sharded ? (td.n_receipts ? (uint64_t)((__uint128_t)nraw_total * lo / td.n_receipts) : 0) : 0;
        const uint64_t ghi = sharded ? (td.n_receipts ? (uint64_t)((__uint128_t)nraw_total * hi / td.n_receipts) : 0) : nraw_total;
        if (plan.nraw != ghi - glo) return fail("share size", plan.nraw, ghi - glo);
        for (uint64_t k = 0; k < plan.nraw; k++) {
            uint8_t c[38];
            for (int b = 0; b < 6; b++) c[b] = (uint8_t)(vals[k].w[4] >> (8 * b));
            memcpy(c + 6, vals[k].w, 32);
            if (memcmp(c, raw38.data() + 38 * (glo + k), 38)) return fail("raw execution list differs at", k, rank);
        }
        // ---- recorded blocks == the oracle's witness for a spec that matches nothing
        ipcfp_event_spec spec;
        memset(&spec, 0, sizeof spec);
        spec.event_signature = "NoSuchEvent(uint256)";
        spec.topic_1 = "nobody";
        ipcfp_event_result* er = nullptr;
        ipcfp_status st = sharded ? oracle_generate_event_proof_shard(os, &td, &spec, lo, hi, world, rank, 0, 1, &er) : oracle_generate_event_proof(os, &td, &spec, 0, 1, &er);
        if (st != IPCFP_OK) return fail("oracle generate failed", (uint64_t)(int64_t)st);
        if (er->n_matching) return fail("the no-match spec matched");
        std::set<std::vector<uint8_t>> got, exp;
        for (uint64_t i = 0; i < n; i++) if (wbits[i >> 5] >> (i & 31) & 1) got.insert(std::vector<uint8_t>(cids + 38 * i, cids + 38 * i + 38));
        for (uint64_t i = 0; i < er->witness.n_blocks; i++) exp.insert(std::vector<uint8_t>(er->witness.cids + 38 * i, er->witness.cids + 38 * i + 38));
        {   // ---- the GENERAL walk over the same share must give the same list and the same recorded blocks
            std::vector<uint32_t> wb2((n + 31) / 32 + 8, 0);
            auto mark2 = [&](const uint8_t* cid) { int32_t b = store_lookup_host_cid(sv, cid); if (b >= 0) witness_mark(wb2.data(), (uint32_t)b); };
            for (uint32_t b = 0; b < P; b++) { mark2(td.parent_cids + 38 * b); mark2(td.parent_txmeta_cids + 38 * b); }
            mark2(td.child_cid); mark2(td.receipts_root);
            for (uint32_t k = 0; k < namt; k++) witness_mark(wb2.data(), f_blk[k]);
            unsigned long long gerr = IPCFP_NO_ERROR;
            uint32_t last_round = 0;
            for (uint32_t k = 0; k < namt; k++) last_round = std::max(last_round, heights[k]);
            std::vector<RawCid> gvals;
            uint64_t gn = 0;
            host_general_walk(sv, namt, f_blk, f_meta, last_round, rlo.data(), rhi.data(), 1, wb2.data(), &gerr, 4 * n + 1024, gvals, gn);
            if (gerr != IPCFP_NO_ERROR) return fail("the general walk reported an error on a well-formed tipset", rank, world);
            if (gn != plan.nraw) return fail("general walk: share size", gn, plan.nraw);
            for (uint64_t k = 0; k < gn; k++) if (memcmp(gvals[k].w, vals[k].w, 40)) return fail("general walk: raw list differs from the dense walk at", k, rank);
            if (wb2 != wbits) return fail("general walk: recorded blocks differ from the dense walk", rank, world);
        }
        if (got != exp) {
            rc = fail("recorded block set differs from the oracle's witness", got.size(), exp.size());
            fprintf(stderr, "  world %u rank %u receipts [%llu,%llu) of %llu; Nraw %llu share [%llu,%llu)\n", world, rank, (unsigned long long)lo, (unsigned long long)hi,
                    (unsigned long long)td.n_receipts, (unsigned long long)nraw_total, (unsigned long long)glo, (unsigned long long)ghi);
            for (uint32_t k = 0; k < namt; k++) fprintf(stderr, "  amt %u: count %llu height %u range [%llu,%llu) clipped [%llu,%llu)\n", k, (unsigned long long)counts[k], heights[k],
                    (unsigned long long)rlo[k], (unsigned long long)rhi[k], (unsigned long long)plan.per_amt[2ull * namt + k], (unsigned long long)plan.per_amt[3ull * namt + k]);
            for (auto& c : exp) if (!got.count(c)) {
                int32_t b = store_lookup_host_cid(sv, c.data());
                uint32_t len = 0;
                const uint8_t* p = b >= 0 ? store_block(sv, (uint32_t)b, len) : nullptr;
                fprintf(stderr, "  only in the oracle's witness: block %d (%u bytes):", b, len);
                for (uint32_t k = 0; k < len && k < 24; k++) fprintf(stderr, " %02x", p[k]);
                fprintf(stderr, "\n");
            }
            for (auto& c : got) if (!exp.count(c)) fprintf(stderr, "  only recorded by the emulated walk: digest %02x%02x%02x%02x..\n", c[6], c[7], c[8], c[9]);
        }
        oracle_event_result_free(er);
    }
    oracle_store_destroy(os);
    synth_free(ts);
    return rc;
}

int main(int argc, char** argv) {
    uint64_t cases = argc > 1 ? strtoull(argv[1], nullptr, 10) : 40;
    uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 7;
    uint64_t walked = 0, runs = 0;
    {   // a root.count larger than the tree can hold (8^(height+1)) must never reach the dense walk: on a completely FULL tree every
        // per-node check of amt_item_dense passes while `count` promises values that do not exist (csrc/walk.cuh)
        struct Case { uint32_t h; uint64_t cnt; bool ok; } cases_[] = {{0, 8, true}, {0, 9, false}, {0, 12, false}, {1, 64, true}, {1, 65, false}, {1, 100, false},
                                                                       {2, 512, true}, {2, 513, false}, {20, 1ull << 40, true}};
        for (const Case& c : cases_) {
            uint32_t hh[1] = {c.h};
            uint64_t cc[1] = {c.cnt}, lo_[1] = {0}, hi_[1] = {UINT64_MAX};
            DensePlan pl = make_dense_plan(1, hh, cc, lo_, hi_, 1ull << 42, 1ull << 43, 1 << 20);
            if (pl.ok != c.ok) return fail("make_dense_plan: capacity rule", c.h, c.cnt);
        }
    }
    for (uint64_t c = 0; c < cases; c++) {
        synth_params sp;
        synth_default_params(&sp);
        uint64_t z = (seed + c) * 0x9E3779B97F4A7C15ull;
        auto rnd = [&]() { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return z; };
        static const uint64_t sizes[] = {0, 1, 2, 7, 8, 9, 63, 64, 65, 100, 511, 513, 1000, 4097, 20000};
        sp.seed = seed * 1000 + c;
        sp.n_receipts = sizes[rnd() % 15];
        sp.events_per_receipt = 1;
        sp.match_ppm = 0;
        sp.n_parents = 1 + (uint32_t)(rnd() % 4);
        sp.dup_msgs = (uint32_t)(rnd() % 3);
        sp.with_state_tree = 0;
        sp.threads = 1;
        static const uint32_t worlds[] = {1, 2, 3, 8};
        for (uint32_t w : worlds) { if (run_case(sp, w, &walked)) return 1; runs += w; }
    }
    printf("ok: dense walk == general walk == oracle on the CPU for %llu tipsets, %llu (tipset, shard) runs, %llu AMT nodes walked\n", (unsigned long long)cases,
           (unsigned long long)runs, (unsigned long long)walked);
    return 0;
}
