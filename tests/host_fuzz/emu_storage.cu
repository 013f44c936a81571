// emu_storage.cu — the storage path (K5) executed ON THE CPU (TEST INFRASTRUCTURE, no GPU needed).
//
// csrc/storage.cuh holds the per-proof device functions: `storage_proof_one` (header → StateRoot → actors HAMT → ActorState →
// EVM state → storage root), `read_storage_slot` (shape sniffing A1→A2→A3→B1→B2→C), `hamt_get`, the value decoders. This
// program compiles them for the host and runs them, spec by spec, over a host copy of the store against the oracle
// (`oracle_generate_storage_proofs`): first on the synthetic state tree as built, then with ONE block on some proof's path
// replaced by a mutated copy under the same CID (the engine does not re-hash unless asked to) — values, found flags, CIDs, the
// per-spec recorded block sets, and, when something is broken, the status and the index of the first failing spec must agree.
//
//   nvcc -std=c++17 -O2 -o emu_storage tests/host_fuzz/emu_storage.cu oracle/oracle.cpp synth/synth.cpp -lpthread && ./emu_storage
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>

#include "host_shims.h"

#include "../../ipc_filecoin_proofs_b200/csrc/storage.cuh"
#include "../../oracle/oracle.h"
#include "../../synth/synth.h"
#include "host_store.h"

using namespace ipcfp;

static uint64_t rs;
static uint64_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return rs; }

static int status_of(const Fail& f) {   // csrc/storage.cu throw_storage_error
    switch (f.code) {
        case DC_MISSING: return IPCFP_ERR_MISSING_BLOCK;
        case DC_STATE_MISMATCH: return IPCFP_ERR_STATE_ROOT_MISMATCH;
        case DC_ACTOR_NOT_FOUND: return IPCFP_ERR_ACTOR_NOT_FOUND;
        case DC_UNSUPPORTED: return IPCFP_ERR_UNSUPPORTED;
        default: return IPCFP_ERR_DECODE;
    }
}

struct Blocks {   // a mutable copy of the flat block arrays
    std::vector<uint8_t> cids, blob;
    std::vector<uint64_t> offs;
    std::vector<uint32_t> lens;
    uint64_t n;
};

// runs both sides over `B`; returns 0 when they agree. *touched receives the blocks on the proofs' paths (engine side).
static int compare(const Blocks& B, const ipcfp_tipset_desc& td, const std::vector<ipcfp_storage_spec>& specs, std::vector<uint32_t>* touched, uint64_t* n_ok, uint64_t* n_err) {
    HostStore hs(B.cids.data(), B.offs.data(), B.lens.data(), B.blob.data(), B.blob.size(), B.n);
    const uint64_t n = specs.size();
    std::vector<uint32_t> wbits((B.n + 31) / 32 + 8, 0), rec_list(n * REC_CAP), rec_n(n, 0);
    std::vector<ipcfp_storage_proof> out(n);
    StorageArgs a;
    memset(&a, 0, sizeof a);
    a.store = hs.view; a.child_cid = td.child_cid; a.state_root_json = td.child_parent_state_root; a.specs = specs.data(); a.n = n; a.out = out.data();
    a.rec_list = rec_list.data(); a.rec_n = rec_n.data(); a.wbits = wbits.data();
    int est = IPCFP_OK;
    uint64_t eidx = 0;
    for (uint64_t t = 0; t < n; t++) {           // the kernel runs one thread per spec; the FIRST failing spec is what gets reported
        Recorder rec{rec_list.data() + t * REC_CAP, 0, wbits.data(), false};
        ipcfp_storage_proof q;
        memset(&q, 0, sizeof q);
        Fail f{0, 0};
        if (!storage_proof_one(a, t, rec, q, f)) { if (est == IPCFP_OK) { est = status_of(f); eidx = t; } continue; }
        if (rec.overflow) { if (est == IPCFP_OK) { est = IPCFP_ERR_UNSUPPORTED; eidx = t; } continue; }
        out[t] = q;
        rec_n[t] = rec.n;
    }
    if (touched) { touched->clear(); for (uint64_t t = 0; t < n; t++) for (uint32_t k = 0; k < rec_n[t]; k++) touched->push_back(rec_list[t * REC_CAP + k]); }
    oracle_store* os = oracle_store_create(B.cids.data(), B.offs.data(), B.lens.data(), B.blob.data(), B.n);
    ipcfp_storage_result* res = nullptr;
    int ost = (int)oracle_generate_storage_proofs(os, &td, specs.data(), n, &res);
    int rc = 0;
    if (ost != est) { fprintf(stderr, "EMU MISMATCH: status engine %d (spec %llu) vs oracle %d (spec %llu)\n", est, (unsigned long long)eidx, ost, (unsigned long long)oracle_last_error_index()); rc = 1; }
    else if (ost != IPCFP_OK) {
        if (oracle_last_error_index() != eidx) { fprintf(stderr, "EMU MISMATCH: failing spec engine %llu vs oracle %llu (status %d)\n", (unsigned long long)eidx, (unsigned long long)oracle_last_error_index(), ost); rc = 1; }
        (*n_err)++;
    } else {
        for (uint64_t t = 0; t < n && !rc; t++) {
            const ipcfp_storage_proof& o = res->proofs[t];
            const ipcfp_storage_proof& e = out[t];
            if (o.actor_id != e.actor_id || memcmp(o.actor_state_cid, e.actor_state_cid, 38) || memcmp(o.storage_root, e.storage_root, 38) || memcmp(o.slot, e.slot, 32) ||
                memcmp(o.value, e.value, 32) || o.found != e.found || o.raw_len != e.raw_len) { fprintf(stderr, "EMU MISMATCH: proof %llu differs (found %u/%u raw_len %u/%u)\n", (unsigned long long)t, e.found, o.found, e.raw_len, o.raw_len); rc = 1; break; }
            std::set<std::vector<uint8_t>> got, exp;
            for (uint32_t k = 0; k < rec_n[t]; k++) { const uint8_t* c = B.cids.data() + 38ull * rec_list[t * REC_CAP + k]; got.insert(std::vector<uint8_t>(c, c + 38)); }
            for (uint64_t k = res->spec_witness_offsets[t]; k < res->spec_witness_offsets[t + 1]; k++) { const uint8_t* c = res->witness.cids + 38ull * res->spec_witness_index[k]; exp.insert(std::vector<uint8_t>(c, c + 38)); }
            if (got != exp) { fprintf(stderr, "EMU MISMATCH: recorded blocks of proof %llu differ (%zu vs %zu)\n", (unsigned long long)t, got.size(), exp.size()); rc = 1; }
        }
        (*n_ok)++;
    }
    if (res) oracle_storage_result_free(res);
    oracle_store_destroy(os);
    return rc;
}

int main(int argc, char** argv) {
    uint64_t cases = argc > 1 ? strtoull(argv[1], nullptr, 10) : 6;
    uint64_t muts = argc > 2 ? strtoull(argv[2], nullptr, 10) : 150;
    rs = argc > 3 ? strtoull(argv[3], nullptr, 10) : 0x51073ull;
    uint64_t n_ok = 0, n_err = 0;
    {   // the hashes read __constant__ tables: make sure the host build sees their values (HAMT paths depend on SHA-256)
        const uint8_t msg[3] = {'a', 'b', 'c'};
        uint32_t h[8];
        sha256(msg, 3, h);
        uint8_t ref[32];
        oracle_sha256(msg, 3, ref);
        for (int i = 0; i < 8; i++) {
            uint32_t be = ((uint32_t)ref[4 * i] << 24) | ((uint32_t)ref[4 * i + 1] << 16) | ((uint32_t)ref[4 * i + 2] << 8) | ref[4 * i + 3];
            if (h[i] != be) { fprintf(stderr, "host build of sha256 is broken (word %d)\n", i); return 2; }
        }
    }
    for (uint64_t c = 0; c < cases; c++) {
        synth_params sp;
        synth_default_params(&sp);
        sp.seed = 900 + c;
        sp.n_receipts = 8;
        sp.events_per_receipt = 1;
        sp.with_state_tree = 1;
        sp.n_actors = 40 + (uint32_t)(rnd() % 2000);
        sp.hamt_entries = 3 + rnd() % 5000;
        sp.threads = 1;
        synth_tipset* ts = synth_build(&sp);
        Blocks B;
        B.n = synth_n_blocks(ts);
        B.cids.assign(synth_cids(ts), synth_cids(ts) + 38 * B.n);
        B.offs.assign(synth_offsets(ts), synth_offsets(ts) + B.n);
        B.lens.assign(synth_lengths(ts), synth_lengths(ts) + B.n);
        B.blob.assign(synth_blob(ts), synth_blob(ts) + synth_blob_size(ts));
        ipcfp_tipset_desc td;
        memset(&td, 0, sizeof td);
        td.parent_epoch = synth_parent_epoch(ts); td.child_epoch = synth_child_epoch(ts); td.n_parents = synth_n_parents(ts);
        td.parent_cids = synth_parent_cids(ts); td.parent_txmeta_cids = synth_parent_txmeta_cids(ts); td.child_cid = synth_child_cid(ts);
        td.receipts_root = synth_receipts_root(ts); td.child_parent_state_root = synth_parent_state_root(ts); td.n_receipts = synth_n_receipts(ts);
        td.events_roots = synth_events_roots(ts); td.has_events_root = synth_has_events_root(ts);
        // specs: the six EVM actors (one per sniffing shape) × present / absent / special slots, + an actor that does not exist
        std::vector<ipcfp_storage_spec> specs;
        for (uint64_t actor : {1001ull, 1002ull, 1003ull, 1004ull, 1005ull, 1006ull}) {
            for (int k = 0; k < 5; k++) {
                ipcfp_storage_spec s;
                memset(&s, 0, sizeof s);
                s.actor_id = actor;
                uint8_t key[32], val[32];
                if (k < 3) synth_storage_entry(ts, k == 2 ? rnd() % sp.hamt_entries : (uint64_t)k, key, val);
                else if (k == 3) synth_storage_absent_key(ts, rnd() % 1000, key);
                else { memset(key, 0, 32); memcpy(key, "calib-subnet-1", 14); }
                oracle_compute_mapping_slot(key, 0, s.slot);
                specs.push_back(s);
            }
        }
        std::vector<uint32_t> touched;
        if (compare(B, td, specs, &touched, &n_ok, &n_err)) return 1;
        {   // an actor that is not in the state tree: both sides must fail at that spec with ACTOR_NOT_FOUND
            std::vector<ipcfp_storage_spec> s2(specs.begin(), specs.begin() + 3);
            s2[1].actor_id = 999999;
            if (compare(B, td, s2, nullptr, &n_ok, &n_err)) return 1;
        }
        std::sort(touched.begin(), touched.end());
        touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
        for (uint64_t m = 0; m < muts; m++) {
            Blocks M = B;
            uint32_t victim = touched[rnd() % touched.size()];
            std::vector<uint8_t> blk(B.blob.begin() + (long)B.offs[victim], B.blob.begin() + (long)B.offs[victim] + B.lens[victim]);
            unsigned nm = 1 + (unsigned)(rnd() % 2);
            for (unsigned k = 0; k < nm; k++) {
                size_t at = rnd() % blk.size();
                switch (rnd() % 5) {
                    case 0: blk[at] = (uint8_t)rnd(); break;
                    case 1: blk[at] ^= (uint8_t)(1u << (rnd() % 8)); break;
                    case 2: blk.erase(blk.begin() + (long)at); break;
                    case 3: blk.insert(blk.begin() + (long)at, (uint8_t)rnd()); break;
                    default: blk.resize(at); break;                       // truncated block
                }
                if (blk.empty()) blk.push_back(0x80);
            }
            while (M.blob.size() % 16) M.blob.push_back(0);
            M.offs[victim] = M.blob.size();
            M.lens[victim] = (uint32_t)blk.size();
            M.blob.insert(M.blob.end(), blk.begin(), blk.end());
            if (rnd() % 16 == 0) {                                        // or the block is simply not there
                M.cids[38ull * victim + 20] ^= 0x5a;
            }
            if (compare(M, td, specs, nullptr, &n_ok, &n_err)) { fprintf(stderr, "  (tipset %llu, mutation %llu of block %u)\n", (unsigned long long)c, (unsigned long long)m, victim); return 1; }
        }
        synth_free(ts);
    }
    printf("ok: storage path on the CPU == oracle for %llu state trees: %llu runs with all proofs equal, %llu runs failing identically\n", (unsigned long long)cases,
           (unsigned long long)n_ok, (unsigned long long)n_err);
    return 0;
}
