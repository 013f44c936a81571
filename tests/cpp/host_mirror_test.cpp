// host_mirror_test.cpp — tests of include/ipcfp.hpp, the C++ host side with the reference's names, written the way tests of the
// reference's own crate would read: build a tipset pair, call generate_event_proof / generate_storage_proof / generate_proof_bundle,
// verify the bundle, tamper with it. TEST CODE: the expected values come from the CPU oracle (oracle/, linked as the checker) on
// the same synthetic tipsets (synth/), converted to the reference's structs by the very same conversion functions.
//
//   host_mirror_test cpu     no device needed: Cid / hex / TipsetDesc / conversions; device calls must fail with IPCFP_ERR_NO_DEVICE
//                            (or succeed when a GPU happens to be present)
//   host_mirror_test gpu     the generators and verifiers on cuda:0 against the oracle
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/ipcfp.hpp"
#include "../../oracle/oracle.h"
#include "../../synth/synth.h"

using namespace ipcfp::host;

static int g_checks = 0;
#define REQUIRE(cond)                                                                      \
    do {                                                                                   \
        g_checks++;                                                                        \
        if (!(cond)) { fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); exit(1); } \
    } while (0)

template <class F>
static ipcfp_status status_of(F&& f) {
    try { f(); } catch (const Error& e) { return e.status; }
    return IPCFP_OK;
}

// ------------------------------------------------------------------------------------------ a synthetic tipset pair as Lotus would serve it
struct Fixture {
    synth_tipset* ts = nullptr;
    oracle_store* os = nullptr;
    ApiTipset parent, child;
    std::vector<ApiReceipt> receipts;
    ipcfp_tipset_desc raw;   // the descriptor straight from the synthetic builder (what the Python tests pass)

    explicit Fixture(const synth_params& p) {
        ts = synth_build(&p);
        REQUIRE(ts != nullptr);
        os = oracle_store_create(synth_cids(ts), synth_offsets(ts), synth_lengths(ts), synth_blob(ts), synth_n_blocks(ts));
        const uint32_t P = synth_n_parents(ts);
        parent.height = synth_parent_epoch(ts);
        child.height = synth_child_epoch(ts);
        for (uint32_t i = 0; i < P; i++) {
            parent.cids.push_back({Cid::from_bytes(synth_parent_cids(ts) + 38 * i).to_string()});
            ApiBlockHeader h;
            h.messages = {Cid::from_bytes(synth_parent_txmeta_cids(ts) + 38 * i).to_string()};
            h.height = parent.height;
            parent.blocks.push_back(h);
        }
        child.cids.push_back({Cid::from_bytes(synth_child_cid(ts)).to_string()});
        ApiBlockHeader ch;
        ch.parent_message_receipts = {Cid::from_bytes(synth_receipts_root(ts)).to_string()};
        ch.parent_state_root = {Cid::from_bytes(synth_parent_state_root(ts)).to_string()};
        ch.height = child.height;
        for (const auto& c : parent.cids) ch.parents.push_back(c);
        child.blocks.push_back(ch);
        const uint64_t n = synth_n_receipts(ts);
        receipts.resize(n);
        for (uint64_t i = 0; i < n; i++)
            if (synth_has_events_root(ts)[i]) receipts[i].events_root = CIDMap{Cid::from_bytes(synth_events_roots(ts) + 38 * i).to_string()};
        memset(&raw, 0, sizeof raw);
        raw.parent_epoch = parent.height; raw.child_epoch = child.height; raw.n_parents = P;
        raw.parent_cids = synth_parent_cids(ts); raw.parent_txmeta_cids = synth_parent_txmeta_cids(ts); raw.child_cid = synth_child_cid(ts);
        raw.receipts_root = synth_receipts_root(ts); raw.child_parent_state_root = synth_parent_state_root(ts);
        raw.n_receipts = n; raw.events_roots = synth_events_roots(ts); raw.has_events_root = synth_has_events_root(ts);
    }
    ~Fixture() { oracle_store_destroy(os); synth_free(ts); }
    GpuBlockstore store(bool verify = true) const {
        return GpuBlockstore::from_flat(synth_cids(ts), synth_offsets(ts), synth_lengths(ts), synth_blob(ts), synth_blob_size(ts), synth_n_blocks(ts), 0, verify);
    }
    std::optional<uint64_t> actor_filter(const synth_params& p) const { return p.has_actor_filter ? std::optional<uint64_t>(synth_target_actor(ts)) : std::nullopt; }
};

static synth_params config(int id) {   // synth/__init__.py::config_params
    synth_params p;
    synth_default_params(&p);
    p.seed = 0x1FC0FFEEull ^ (uint64_t)id;
    if (id == 1) { p.n_receipts = 64; p.events_per_receipt = 8; p.match_ppm = 125000; p.has_actor_filter = 0; p.same_topic1 = 1; p.bw3_permille = 0; p.dup_msgs = 2; }
    else if (id == 2) { p.n_receipts = 10000; p.events_per_receipt = 8; p.match_ppm = 10000; p.has_actor_filter = 1; p.bw3_permille = 100; }
    else { p.n_receipts = 64; p.events_per_receipt = 8; p.match_ppm = 20000; p.with_state_tree = 1; p.hamt_entries = 20000; p.n_actors = 2048; }   // configs[2], small HAMT
    return p;
}

// the oracle's answer, in the reference's structs (same conversion as the engine's result goes through)
static EventProofBundle oracle_event_bundle(const Fixture& f, const std::string& sig, const std::string& t1, std::optional<uint64_t> filter) {
    TipsetDesc t(f.parent, f.child, f.receipts);
    ipcfp_event_spec spec = spec_c(sig, t1, filter);
    ipcfp_event_result* r = nullptr;
    REQUIRE(oracle_generate_event_proof(f.os, t.c(), &spec, 0, 1, &r) == IPCFP_OK);
    EventProofBundle b;
    b.proofs = event_proofs(*r, t);
    b.blocks = proof_blocks(r->witness);
    oracle_event_result_free(r);
    return b;
}

// ------------------------------------------------------------------------------------------ cpu
static int run_cpu() {
    // Cid <-> string: three constants of the public Filecoin chain (tests/test_oracle_cpu.py)
    for (const char* s : {"bafy2bzacecmda75ovposbdateg7eyhwij65zklgyijgcjwynlklmqazpwlhba", "bafy2bzacedijw74yui7otvo63nfl3hdq2vdzuy7wx2tnptwed6zml4vvz7wee",
                          "bafy2bzaceamp42wmmgr2g2ymg46euououzfyck7szknvfacqscohrvaikwfay"}) {
        Cid c = Cid::try_from(s);
        const uint8_t pre[6] = {0x01, 0x71, 0xa0, 0xe4, 0x02, 0x20};
        REQUIRE(memcmp(c.bytes.data(), pre, 6) == 0);
        REQUIRE(c.to_string() == s);
    }
    {   // the empty HAMT node 82 40 80 hashes to the third constant
        const uint8_t node[3] = {0x82, 0x40, 0x80};
        Cid c = Cid::try_from("bafy2bzaceamp42wmmgr2g2ymg46euououzfyck7szknvfacqscohrvaikwfay");
        uint8_t d[32];
        oracle_blake2b256(node, 3, d);
        REQUIRE(memcmp(c.bytes.data() + 6, d, 32) == 0);
    }
    REQUIRE(status_of([] { Cid::try_from("bafy2bzace"); }) == IPCFP_ERR_INVALID_ARG);
    REQUIRE(status_of([] { Cid::try_from("Bafy2bzacecmda75ovposbdateg7eyhwij65zklgyijgcjwynlklmqazpwlhba"); }) == IPCFP_ERR_INVALID_ARG);
    REQUIRE(status_of([] { Cid::try_from("bafy2bzacecmda75ovposbdateg7eyhwij65zklgyijgcjwynlklmqazpwlhb1"); }) == IPCFP_ERR_INVALID_ARG);
    REQUIRE(status_of([] { Cid::try_from("bafy2bzacecmda75ovposbdateg7eyhwij65zklgyijgcjwynlklmqazpwlhbb"); }) == IPCFP_ERR_INVALID_ARG);   // non-zero trailing bits

    // `Ord` of Cid == the oracle's sort (cid 0.11: version, codec, multihash code, size, digest), on CIDs of several codecs / hash codes
    {
        uint64_t z = 0x243F6A8885A308D3ull;
        auto rnd = [&]() { z ^= z << 13; z ^= z >> 7; z ^= z << 17; return z; };
        const uint8_t codecs[3] = {0x71, 0x55, 0x70};
        const uint8_t codes[3][3] = {{0xa0, 0xe4, 0x02}, {0xc0, 0xe4, 0x02}, {0x80, 0x80, 0x01}};
        std::vector<Cid> v;
        std::vector<uint8_t> flat;
        for (int i = 0; i < 3000; i++) {
            Cid c;
            c.bytes[0] = 0x01; c.bytes[1] = codecs[rnd() % 3];
            memcpy(c.bytes.data() + 2, codes[rnd() % 3], 3);
            c.bytes[5] = 0x20;
            for (int k = 6; k < 38; k++) c.bytes[k] = (uint8_t)(rnd() % 4 ? rnd() : 0);   // shared prefixes and zero runs
            if (i % 50 == 1) c = v[(size_t)(rnd() % v.size())];                            // duplicates
            v.push_back(c);
            flat.insert(flat.end(), c.bytes.begin(), c.bytes.end());
        }
        const uint64_t n = oracle_sort_unique_cids(flat.data(), v.size());
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        REQUIRE(v.size() == n);
        for (uint64_t i = 0; i < n; i++) REQUIRE(memcmp(v[i].bytes.data(), flat.data() + 38 * i, 38) == 0);
    }

    // parse_cid / parse_cids / create_event_filter (common/witness.rs:60-72, events/verifier.rs:28-41)
    {
        const std::string good = "bafy2bzaceamp42wmmgr2g2ymg46euououzfyck7szknvfacqscohrvaikwfay";
        REQUIRE(parse_cid(good, "child block").to_string() == good);
        REQUIRE(parse_cids({good, good}, "parent tipset").size() == 2);
        bool named = false;
        try { parse_cid("nonsense", "child block"); } catch (const Error& e) { named = e.status == IPCFP_ERR_INVALID_ARG && std::string(e.what()).find("child block") != std::string::npos; }
        REQUIRE(named);
        EventProofSpec f = create_event_filter("NewTopDownMessage(bytes32,uint256)", "calib-subnet-1");
        REQUIRE(f.event_signature == "NewTopDownMessage(bytes32,uint256)" && f.topic_1 == "calib-subnet-1" && !f.actor_id_filter);
    }

    // hex / padding helpers (common/evm.rs:72-100)
    {
        const uint8_t b[3] = {0x00, 0xab, 0xff};
        REQUIRE(to_hex0x(b, 3) == "0x00abff");
        REQUIRE(to_hex0x(b, 0) == "0x");
        REQUIRE(from_hex0x("0x00ABff") == std::vector<uint8_t>(b, b + 3));
        REQUIRE(from_hex0x("0x").empty());
        REQUIRE(status_of([] { from_hex0x("0x0"); }) == IPCFP_ERR_INVALID_ARG);
        REQUIRE(status_of([] { from_hex0x("0xzz"); }) == IPCFP_ERR_INVALID_ARG);
        REQUIRE(status_of([] { from_hex32("0x00"); }) == IPCFP_ERR_INVALID_ARG);
        H256 t = ascii_to_bytes32("calib-subnet-1");
        REQUIRE(memcmp(t.data(), "calib-subnet-1", 14) == 0 && t[14] == 0 && t[31] == 0);
        REQUIRE(ascii_to_bytes32(std::string(40, 'x'))[31] == 'x');
        std::vector<uint8_t> longv(40);
        for (int i = 0; i < 40; i++) longv[i] = (uint8_t)i;
        H256 l = left_pad_32(longv);
        REQUIRE(l[0] == 8 && l[31] == 39);   // longer than 32: the LAST 32 bytes
        H256 s = left_pad_32({1, 2});
        REQUIRE(s[29] == 0 && s[30] == 1 && s[31] == 2);
        REQUIRE(left_pad_32({}) == H256{});
    }

    // TipsetDesc packs (parent, child, receipts) into exactly the descriptor the synthetic builder hands the C ABI
    synth_params p1 = config(1);
    Fixture f(p1);
    {
        TipsetDesc t(f.parent, f.child, f.receipts);
        const ipcfp_tipset_desc* d = t.c();
        REQUIRE(d->parent_epoch == f.raw.parent_epoch && d->child_epoch == f.raw.child_epoch && d->n_parents == f.raw.n_parents && d->n_receipts == f.raw.n_receipts);
        REQUIRE(memcmp(d->parent_cids, f.raw.parent_cids, 38 * d->n_parents) == 0);
        REQUIRE(memcmp(d->parent_txmeta_cids, f.raw.parent_txmeta_cids, 38 * d->n_parents) == 0);
        REQUIRE(memcmp(d->child_cid, f.raw.child_cid, 38) == 0 && memcmp(d->receipts_root, f.raw.receipts_root, 38) == 0);
        REQUIRE(memcmp(d->child_parent_state_root, f.raw.child_parent_state_root, 38) == 0);
        REQUIRE(memcmp(d->has_events_root, f.raw.has_events_root, d->n_receipts) == 0);
        for (uint64_t i = 0; i < d->n_receipts; i++)
            if (d->has_events_root[i]) REQUIRE(memcmp(d->events_roots + 38 * i, f.raw.events_roots + 38 * i, 38) == 0);
        ApiTipset no_child = f.child;
        no_child.cids.clear();
        REQUIRE(status_of([&] { TipsetDesc bad(f.parent, no_child, f.receipts); }) == IPCFP_ERR_INVALID_ARG);
    }

    // the conversions on an oracle result: the reference's EventProof fields (events/generator.rs:276-296)
    {
        EventProofBundle b = oracle_event_bundle(f, synth_event_signature(f.ts), synth_topic1(f.ts), f.actor_filter(p1));
        REQUIRE(!b.proofs.empty() && b.proofs.size() == synth_n_selected(f.ts));
        for (size_t i = 0; i < b.proofs.size(); i++) {
            const EventProof& e = b.proofs[i];
            REQUIRE(e.parent_epoch == f.parent.height && e.child_epoch == f.child.height && e.child_block_cid == f.child.cids[0].cid);
            REQUIRE(e.parent_tipset_cids.size() == f.parent.cids.size() && e.exec_index == synth_selected(f.ts)[i]);
            REQUIRE(Cid::try_from(e.message_cid).to_string() == e.message_cid);
            REQUIRE(e.event_data.topics.size() >= 2);
            for (const auto& t : e.event_data.topics) REQUIRE(t.size() == 66 && t.compare(0, 2, "0x") == 0);
            H256 t1 = ascii_to_bytes32(synth_topic1(f.ts));
            REQUIRE(e.event_data.topics[1] == to_hex0x(t1.data(), 32));
        }
        REQUIRE(std::is_sorted(b.blocks.begin(), b.blocks.end(), [](const ProofBlock& x, const ProofBlock& y) { return x.cid < y.cid; }));
        for (const auto& blk : b.blocks) {   // every witness block hashes to its CID
            uint8_t d[32];
            oracle_blake2b256(blk.data.data(), blk.data.size(), d);
            REQUIRE(memcmp(blk.cid.bytes.data() + 6, d, 32) == 0);
        }
    }

    // wire format: to_json == the C ABI's rendering of the same (oracle) result byte for byte; bundle_from_json brings every field back
    {
        TipsetDesc t(f.parent, f.child, f.receipts);
        ipcfp_event_spec spec = spec_c(synth_event_signature(f.ts), synth_topic1(f.ts), f.actor_filter(p1));
        ipcfp_event_result* r = nullptr;
        REQUIRE(oracle_generate_event_proof(f.os, t.c(), &spec, 0, 1, &r) == IPCFP_OK);
        EventProofBundle b;
        b.proofs = event_proofs(*r, t);
        b.blocks = proof_blocks(r->witness);
        char* text = nullptr;
        uint64_t len = 0;
        REQUIRE(ipcfp_event_result_to_json(r, t.c(), &text, &len) == IPCFP_OK);
        oracle_event_result_free(r);
        const std::string mine = to_json(b);
        REQUIRE(mine.size() == len && mine == std::string(text, len));
        ipcfp_json_free(text);
        UnifiedProofBundle back = bundle_from_json(mine);
        REQUIRE(back.storage_proofs.empty() && back.event_proofs.size() == b.proofs.size() && back.blocks.size() == b.blocks.size());
        for (size_t i = 0; i < b.proofs.size(); i++) REQUIRE(back.event_proofs[i] == b.proofs[i]);
        for (size_t i = 0; i < b.blocks.size(); i++) REQUIRE(back.blocks[i] == b.blocks[i]);
        EventProofBundle again;
        again.proofs = back.event_proofs;
        again.blocks = back.blocks;
        REQUIRE(to_json(again) == mine);
        REQUIRE(status_of([&] { bundle_from_json(mine.substr(0, mine.size() - 1)); }) == IPCFP_ERR_INVALID_ARG);
        REQUIRE(status_of([&] { bundle_from_json(mine + "x"); }) == IPCFP_ERR_INVALID_ARG);   // trailing characters
        std::string esc;
        detail::json_string(esc, std::string("a\"b\\c\n\x01", 7));
        REQUIRE(esc == "\"a\\\"b\\\\c\\n\\u0001\"");
    }
    {
        synth_params p3 = config(3);
        Fixture g(p3);
        TipsetDesc t(g.parent, g.child, g.receipts);
        H256 k = ascii_to_bytes32("calib-subnet-1"), special;
        oracle_compute_mapping_slot(k.data(), 0, special.data());
        ipcfp_storage_spec cs[2];
        memset(cs, 0, sizeof cs);
        cs[0].actor_id = 1001; cs[1].actor_id = 1003;
        memcpy(cs[0].slot, special.data(), 32); memcpy(cs[1].slot, special.data(), 32);
        const std::string sig = synth_event_signature(g.ts), t1 = synth_topic1(g.ts);
        ipcfp_event_spec ce[2] = {spec_c(sig, t1, g.actor_filter(p3)), spec_c(sig, "calib-subnet-2", std::nullopt)};
        ipcfp_bundle* ob = nullptr;
        REQUIRE(oracle_generate_proof_bundle(g.os, t.c(), cs, 2, ce, 2, &ob) == IPCFP_OK);
        UnifiedProofBundle u;
        for (uint64_t i = 0; i < ob->storage->n_proofs; i++) u.storage_proofs.push_back(storage_proof(ob->storage->proofs[i], t));
        for (uint64_t q = 0; q < ob->n_event_results; q++) { auto e = event_proofs(*ob->events[q], t); u.event_proofs.insert(u.event_proofs.end(), e.begin(), e.end()); }
        u.blocks = proof_blocks(ob->witness);
        char* text = nullptr;
        uint64_t len = 0;
        REQUIRE(ipcfp_bundle_to_json(ob, t.c(), &text, &len) == IPCFP_OK);
        oracle_bundle_free(ob);
        const std::string mine = to_json(u);
        REQUIRE(mine == std::string(text, len) && u.storage_proofs.size() == 2 && !u.event_proofs.empty());
        ipcfp_json_free(text);
        UnifiedProofBundle back = bundle_from_json(mine);
        REQUIRE(back.storage_proofs.size() == 2 && back.storage_proofs[0] == u.storage_proofs[0] && back.storage_proofs[1] == u.storage_proofs[1]);
        REQUIRE(back.event_proofs.size() == u.event_proofs.size() && back.blocks.size() == u.blocks.size());
        for (size_t i = 0; i < u.event_proofs.size(); i++) REQUIRE(back.event_proofs[i] == u.event_proofs[i]);
        for (size_t i = 0; i < u.blocks.size(); i++) REQUIRE(back.blocks[i] == u.blocks[i]);
        REQUIRE(to_json(back) == mine);
    }

    // no CPU path: without a device the store cannot be created (with one, it can — then this is simply a second smoke test)
    ipcfp_status st = status_of([&] { GpuBlockstore s = f.store(); REQUIRE(s.n_blocks() == synth_n_blocks(f.ts)); });
    REQUIRE(st == IPCFP_ERR_NO_DEVICE || st == IPCFP_OK);
    if (st == IPCFP_ERR_NO_DEVICE) {
        REQUIRE(status_of([] { compute_mapping_slot(H256{}, 0); }) == IPCFP_ERR_NO_DEVICE);
        EventProofBundle b = oracle_event_bundle(f, synth_event_signature(f.ts), synth_topic1(f.ts), std::nullopt);
        auto yes_ts = [](int64_t, const std::vector<Cid>&) { return true; };
        auto yes_h = [](int64_t, const Cid&) { return true; };
        REQUIRE(status_of([&] { verify_event_proof(b, yes_ts, yes_h); }) == IPCFP_ERR_NO_DEVICE);
        // … but a proof whose anchors are not trusted is rejected on the host before any device work (verify_trust_anchors)
        auto no_h = [](int64_t, const Cid&) { return false; };
        StorageProof sp;
        sp.child_block_cid = f.child.cids[0].cid;
        REQUIRE(verify_storage_proof(sp, {}, no_h) == false);
    }
    {   // the multi-GPU surface compiles, links and fails cleanly where NCCL / a device is missing (run for real by tests/test_parallel.py)
        auto fp = &generate_event_proof_sharded;
        REQUIRE(fp != nullptr);
        const ipcfp_status ids = status_of([] { (void)ShardedComm::unique_id(); });
        REQUIRE(ids == IPCFP_OK || ids == IPCFP_ERR_NCCL || ids == IPCFP_ERR_NO_DEVICE);
    }
    printf("ok: cpu checks of include/ipcfp.hpp (%d assertions)%s\n", g_checks, st == IPCFP_OK ? " [a CUDA device was present]" : "");
    return 0;
}

// ------------------------------------------------------------------------------------------ gpu
static void expect_equal(const EventProofBundle& got, const EventProofBundle& exp) {
    REQUIRE(got.proofs.size() == exp.proofs.size());
    for (size_t i = 0; i < got.proofs.size(); i++) REQUIRE(got.proofs[i] == exp.proofs[i]);
    REQUIRE(got.blocks.size() == exp.blocks.size());
    for (size_t i = 0; i < got.blocks.size(); i++) REQUIRE(got.blocks[i] == exp.blocks[i]);
}

static int run_gpu() {
    auto yes_ts = [](int64_t, const std::vector<Cid>&) { return true; };
    auto yes_h = [](int64_t, const Cid&) { return true; };
    auto no_ts = [](int64_t, const std::vector<Cid>&) { return false; };
    auto no_h = [](int64_t, const Cid&) { return false; };
    auto all = [](const std::vector<bool>& v, bool want) { return std::all_of(v.begin(), v.end(), [&](bool b) { return b == want; }); };

    // ---- generate_event_proof + verify_event_proof, configs[0] and configs[1]
    for (int id : {1, 2}) {
        synth_params p = config(id);
        Fixture f(p);
        GpuBlockstore store = f.store();
        const std::string sig = synth_event_signature(f.ts), t1 = synth_topic1(f.ts);
        EventProofBundle got = generate_event_proof(store, f.parent, f.child, f.receipts, sig, t1, f.actor_filter(p));
        EventProofBundle exp = oracle_event_bundle(f, sig, t1, f.actor_filter(p));
        REQUIRE(!exp.proofs.empty());
        expect_equal(got, exp);

        // Blockstore: get / has / put_keyed
        const ProofBlock& b0 = got.blocks[got.blocks.size() / 2];
        auto bytes = store.get(b0.cid);
        REQUIRE(bytes.has_value() && *bytes == b0.data && store.has(b0.cid));
        Cid unknown = b0.cid;
        unknown.bytes[37] ^= 1;
        REQUIRE(!store.get(unknown).has_value() && !store.has(unknown));
        bool threw = false;
        try { store.put_keyed(b0.cid, b0.data); } catch (const std::logic_error&) { threw = true; }
        REQUIRE(threw);

        // verification: every proof holds; untrusted anchors reject on the host; check_event = the spec accepts, another subnet rejects
        REQUIRE(all(verify_event_proof(got, yes_ts, yes_h), true) && !got.proofs.empty());
        REQUIRE(all(verify_event_proof(got, no_ts, yes_h), false));
        REQUIRE(all(verify_event_proof(got, yes_ts, no_h), false));
        EventProofSpec same{sig, t1, f.actor_filter(p)}, other{sig, "some-other-subnet", std::nullopt};
        REQUIRE(all(verify_event_proof(got, yes_ts, yes_h, &same), true));
        REQUIRE(all(verify_event_proof(got, yes_ts, yes_h, &other), false));
        // a forged claim: another event index
        EventProofBundle forged = got;
        forged.proofs[0].event_index += 1;
        std::vector<bool> fr;
        ipcfp_status fst = status_of([&] { fr = verify_event_proof(forged, yes_ts, yes_h); });
        REQUIRE(fst != IPCFP_OK || fr[0] == false);
        // a witness block damaged under its CID is caught by the Blake2b check the reference's load_witness_store leaves out
        EventProofBundle damaged = got;
        damaged.blocks[3].data[damaged.blocks[3].data.size() / 2] ^= 0x40;
        REQUIRE(status_of([&] { verify_event_proof(damaged, yes_ts, yes_h); }) == IPCFP_ERR_CID_MISMATCH);
        // a missing witness block: an error or a rejection, never an acceptance of everything
        if (id == 1) {
            for (size_t drop = 0; drop < got.blocks.size(); drop++) {
                EventProofBundle fewer = got;
                fewer.blocks.erase(fewer.blocks.begin() + (long)drop);
                std::vector<bool> r;
                ipcfp_status st = status_of([&] { r = verify_event_proof(fewer, yes_ts, yes_h); });
                REQUIRE(st != IPCFP_OK || !all(r, true));
            }
        }
        // a spec nothing matches: empty proofs, the base witness only
        EventProofBundle none = generate_event_proof(store, f.parent, f.child, f.receipts, "Nothing(uint256)", t1, std::nullopt);
        expect_equal(none, oracle_event_bundle(f, "Nothing(uint256)", t1, std::nullopt));
        REQUIRE(none.proofs.empty() && !none.blocks.empty());
        // a store that lacks a block the scan reads: the reference's `missing …` error — same status, same receipt index as the oracle
        if (id == 1) {
            uint64_t r5 = 5;
            while (r5 < f.receipts.size() && !f.receipts[r5].events_root) r5++;
            REQUIRE(r5 < f.receipts.size());
            const Cid victim = Cid::try_from(f.receipts[r5].events_root->cid);
            std::vector<uint8_t> hc, hb;
            std::vector<uint64_t> ho;
            std::vector<uint32_t> hl;
            for (uint64_t i = 0; i < synth_n_blocks(f.ts); i++) {
                if (Cid::from_bytes(synth_cids(f.ts) + 38 * i) == victim) continue;
                hc.insert(hc.end(), synth_cids(f.ts) + 38 * i, synth_cids(f.ts) + 38 * (i + 1));
                ho.push_back(hb.size());
                hl.push_back(synth_lengths(f.ts)[i]);
                hb.insert(hb.end(), synth_blob(f.ts) + synth_offsets(f.ts)[i], synth_blob(f.ts) + synth_offsets(f.ts)[i] + synth_lengths(f.ts)[i]);
            }
            REQUIRE(ho.size() + 1 <= synth_n_blocks(f.ts));
            GpuBlockstore holed = GpuBlockstore::from_flat(hc.data(), ho.data(), hl.data(), hb.data(), hb.size(), ho.size());
            REQUIRE(!holed.has(victim));
            uint64_t idx = 0;
            ipcfp_status st = IPCFP_OK;
            try { generate_event_proof(holed, f.parent, f.child, f.receipts, sig, t1, f.actor_filter(p)); } catch (const Error& e) { st = e.status; idx = e.index; }
            oracle_store* hos = oracle_store_create(hc.data(), ho.data(), hl.data(), hb.data(), ho.size());
            TipsetDesc t(f.parent, f.child, f.receipts);
            ipcfp_event_spec spec = spec_c(sig, t1, f.actor_filter(p));
            ipcfp_event_result* orr = nullptr;
            const ipcfp_status ost = oracle_generate_event_proof(hos, t.c(), &spec, 0, 1, &orr);
            REQUIRE(ost == IPCFP_ERR_MISSING_BLOCK && orr == nullptr);
            REQUIRE(st == ost && idx == oracle_last_error_index());
            oracle_store_destroy(hos);
        }
    }

    // ---- storage side and the unified bundle, configs[2] with a small HAMT
    {
        synth_params p = config(3);
        Fixture f(p);
        GpuBlockstore store = f.store();
        uint8_t key[32], val[32];
        const uint32_t vlen = synth_storage_entry(f.ts, 77, key, val);
        H256 k77;
        memcpy(k77.data(), key, 32);
        H256 slot77 = compute_mapping_slot(k77, 0);
        uint8_t want[32];
        oracle_compute_mapping_slot(key, 0, want);
        REQUIRE(memcmp(slot77.data(), want, 32) == 0);
        REQUIRE(to_hex0x(compute_mapping_slot(H256{}, 1).data(), 32) == "0xa6eef7e35abe7026729641147f7915573c7e97b47efa546f5f6e3230263bcb49");   // public Solidity vector
        REQUIRE(to_hex0x(hash_event_signature("Transfer(address,address,uint256)").data(), 32) == "0xddf252ad1be2c89b69c2b068fc378daa952ba7f163c4a11628f55a4df523b3ef");
        REQUIRE(to_hex0x(keccak256({}).data(), 32) == "0xc5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470");
        REQUIRE(to_hex0x(keccak256({0x60, 0x80}).data(), 32) == "0x1a578b7a4b0b5755db6d121b4118d4bc68fe170dca840c59bc922f14175a76b0");   // held by the reference tree (forge-std)

        // read_storage_slot: Some(bytes) / None
        const Cid root = Cid::from_bytes(synth_storage_root(f.ts));
        auto v = read_storage_slot(store, root, slot77);
        REQUIRE(v.has_value() && v->size() == vlen && memcmp(v->data(), val, vlen) == 0);
        synth_storage_absent_key(f.ts, 3, key);
        H256 ka;
        memcpy(ka.data(), key, 32);
        REQUIRE(!read_storage_slot(store, root, compute_mapping_slot(ka, 0)).has_value());

        // generate_storage_proof == the oracle's, for a present, the special and an absent slot, over the six root shapes
        const H256 special = calculate_storage_slot("calib-subnet-1", 0);
        for (uint64_t actor : {1001ull, 1002ull, 1003ull, 1004ull, 1005ull, 1006ull}) {
            for (const H256& slot : {slot77, special, compute_mapping_slot(ka, 0)}) {
                auto got = generate_storage_proof(store, f.parent, f.child, actor, slot);
                TipsetDesc t(f.parent, f.child, {});
                ipcfp_storage_spec s;
                memset(&s, 0, sizeof s);
                s.actor_id = actor;
                memcpy(s.slot, slot.data(), 32);
                ipcfp_storage_result* r = nullptr;
                REQUIRE(oracle_generate_storage_proofs(f.os, t.c(), &s, 1, &r) == IPCFP_OK);
                StorageProof exp = storage_proof(r->proofs[0], t);
                std::vector<ProofBlock> expb = proof_blocks(r->witness);
                oracle_storage_result_free(r);
                REQUIRE(got.first == exp);
                REQUIRE(got.second.size() == expb.size());
                for (size_t i = 0; i < expb.size(); i++) REQUIRE(got.second[i] == expb[i]);
                REQUIRE(verify_storage_proof(got.first, got.second, yes_h));
                REQUIRE(!verify_storage_proof(got.first, got.second, no_h));
                StorageProof lie = got.first;
                lie.value[lie.value.size() - 1] = lie.value.back() == '0' ? '1' : '0';
                REQUIRE(!verify_storage_proof(lie, got.second, yes_h));
            }
        }
        // an actor that is not in the state tree: "actor not found" (common/decode.rs:39) — the oracle's status
        {
            TipsetDesc t(f.parent, f.child, {});
            ipcfp_storage_spec s;
            memset(&s, 0, sizeof s);
            s.actor_id = 999999;
            memcpy(s.slot, special.data(), 32);
            ipcfp_storage_result* r = nullptr;
            const ipcfp_status ost = oracle_generate_storage_proofs(f.os, t.c(), &s, 1, &r);
            REQUIRE(ost == IPCFP_ERR_ACTOR_NOT_FOUND);
            REQUIRE(status_of([&] { generate_storage_proof(store, f.parent, f.child, 999999, special); }) == ost);
        }

        // generate_proof_bundle == the oracle's bundle; verify_proof_bundle accepts all of it
        const std::string sig = synth_event_signature(f.ts), t1 = synth_topic1(f.ts);
        std::vector<StorageProofSpec> ss = {{1001, special}, {1003, special}};
        std::vector<EventProofSpec> es = {{sig, t1, f.actor_filter(p)}, {sig, "calib-subnet-2", std::nullopt}};
        UnifiedProofBundle got = generate_proof_bundle(store, f.parent, f.child, f.receipts, ss, es);
        {
            TipsetDesc t(f.parent, f.child, f.receipts);
            std::vector<ipcfp_storage_spec> cs(2);
            for (int i = 0; i < 2; i++) { memset(&cs[i], 0, sizeof cs[i]); cs[i].actor_id = ss[i].actor_id; memcpy(cs[i].slot, ss[i].slot.data(), 32); }
            std::vector<ipcfp_event_spec> ce = {spec_c(es[0].event_signature, es[0].topic_1, es[0].actor_id_filter), spec_c(es[1].event_signature, es[1].topic_1, es[1].actor_id_filter)};
            ipcfp_bundle* ob = nullptr;
            REQUIRE(oracle_generate_proof_bundle(f.os, t.c(), cs.data(), 2, ce.data(), 2, &ob) == IPCFP_OK);
            REQUIRE(got.storage_proofs.size() == ob->storage->n_proofs);
            for (size_t i = 0; i < got.storage_proofs.size(); i++) REQUIRE(got.storage_proofs[i] == storage_proof(ob->storage->proofs[i], t));
            std::vector<EventProof> ep;
            for (uint64_t k = 0; k < ob->n_event_results; k++) { auto e = event_proofs(*ob->events[k], t); ep.insert(ep.end(), e.begin(), e.end()); }
            REQUIRE(got.event_proofs.size() == ep.size() && !ep.empty());
            for (size_t i = 0; i < ep.size(); i++) REQUIRE(got.event_proofs[i] == ep[i]);
            std::vector<ProofBlock> eb = proof_blocks(ob->witness);
            REQUIRE(got.blocks.size() == eb.size());
            for (size_t i = 0; i < eb.size(); i++) REQUIRE(got.blocks[i] == eb[i]);
            oracle_bundle_free(ob);
        }
        UnifiedVerificationResult vr = verify_proof_bundle(got, yes_ts, yes_h);
        REQUIRE(vr.storage_results.size() == 2 && vr.event_results.size() == got.event_proofs.size() && vr.all_valid());
        REQUIRE(!verify_proof_bundle(got, yes_ts, no_h).all_valid());
        // over the wire and back: serde_json text → bundle → verification
        REQUIRE(verify_proof_bundle(bundle_from_json(to_json(got)), yes_ts, yes_h).all_valid());
    }
    printf("ok: include/ipcfp.hpp on cuda:0 == the oracle (%d assertions, %llu kernel launches)\n", g_checks, (unsigned long long)ipcfp_kernel_launch_count());
    return 0;
}

int main(int argc, char** argv) {
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    try {
        return mode == "gpu" ? run_gpu() : run_cpu();
    } catch (const Error& e) {
        fprintf(stderr, "unexpected ipcfp::host::Error status %d index %llu: %s\n", (int)e.status, (unsigned long long)e.index, e.what());
        return 1;
    }
}
