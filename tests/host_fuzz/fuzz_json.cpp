// fuzz_json.cpp — mutation fuzz of ipcfp_bundle_from_json (csrc/bundle_parse.cpp), meant to be built with
// -fsanitize=address,undefined (tests/test_bundle_json.py does). TEST INFRASTRUCTURE ONLY; plain g++, no CUDA, no oracle.
//
//   fuzz_json <mutations per seed> <rng seed> <seed.json>...
//
// Every seed file is a well-formed bundle (rendered by bundle_json.py from oracle results). Each mutation applies 1..4 edits
// (byte flips, token splices from a JSON/hex/base64/base32-relevant dictionary, deletions, duplications of a span, truncation),
// parses the text and — when the parser accepts it — touches every byte the returned object promises: CIDs, offsets / lengths
// inside the witness blob, every proof record, and the topics / data ranges each event proof names inside the data blob.
// A parser that over-reads its input, or returns ranges outside its own buffers, dies under the sanitizer; a range that is merely
// WRONG (outside the blob but still mapped) is caught by the explicit bounds checks below.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ipcfp.h"

static uint64_t rs;
static uint64_t rnd() { uint64_t z = (rs += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

static const char* DICT[] = {"\"", "\\", "{", "}", "[", "]", ",", ":", "null", "true", "false", "0x", "0X", "-", "-0", "1e999", "1.5", "18446744073709551615",
                             "18446744073709551616", "99999999999999999999999", "\\u0000", "\\ud800", "\\udc00", "\\ud83d\\ude00", "\\u00e9", "=", "==", "b", "B",
                             "bafy2bzace", "/", "{\"/\":\"", "\"}", " ", "\n", "\t", "\x01", "\xff", "\xc3\x28", "256", "-1", "38", "[1,113,160,228,2,32", "cid",
                             "data", "blocks", "event_proofs", "storage_proofs", "parent_epoch", "child_epoch", "topics", "0x00", "AAAA", "A", "===="};

static volatile uint64_t sink;
static void touch(const void* p, uint64_t n) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t s = 0;
    for (uint64_t i = 0; i < n; i++) s += b[i];
    sink += s;
}

static int check(const ipcfp_parsed_bundle* pb) {
    const ipcfp_witness& w = pb->witness;
    if (w.n_blocks) {
        touch(w.cids, 38 * w.n_blocks);
        touch(w.offsets, 8 * w.n_blocks);
        touch(w.lengths, 4 * w.n_blocks);
    }
    for (uint64_t i = 0; i < w.n_blocks; i++) {
        if (w.offsets[i] > w.blob_size || w.lengths[i] > w.blob_size - w.offsets[i]) { fprintf(stderr, "block %llu outside the blob\n", (unsigned long long)i); return 1; }
        touch(w.blob + w.offsets[i], w.lengths[i]);
    }
    if (w.blob_size) touch(w.blob, w.blob_size);
    if (pb->n_storage_proofs) touch(pb->storage_proofs, pb->n_storage_proofs * sizeof(ipcfp_storage_proof));
    if (pb->n_event_proofs) touch(pb->event_proofs, pb->n_event_proofs * sizeof(ipcfp_event_proof));
    if (pb->data_blob_size) touch(pb->data_blob, pb->data_blob_size);
    for (uint64_t i = 0; i < pb->n_event_proofs; i++) {
        const ipcfp_event_proof& p = pb->event_proofs[i];
        if (p.topics_off > pb->data_blob_size || 32ull * p.n_topics > pb->data_blob_size - p.topics_off ||
            p.data_off > pb->data_blob_size || p.data_len > pb->data_blob_size - p.data_off) {
            fprintf(stderr, "proof %llu names bytes outside the data blob\n", (unsigned long long)i);
            return 1;
        }
    }
    const ipcfp_tipset_desc& t = pb->tipset;
    if (t.n_parents) { if (!t.parent_cids) { fprintf(stderr, "n_parents without parent_cids\n"); return 1; } touch(t.parent_cids, 38ull * t.n_parents); }
    if (t.child_cid) touch(t.child_cid, 38);
    if (t.child_parent_state_root) touch(t.child_parent_state_root, 38);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: fuzz_json <mutations per seed> <rng seed> <seed.json>...\n"); return 2; }
    const uint64_t muts = strtoull(argv[1], nullptr, 10);
    rs = strtoull(argv[2], nullptr, 10);
    uint64_t accepted = 0, refused = 0, total = 0;
    for (int f = 3; f < argc; f++) {
        FILE* fp = fopen(argv[f], "rb");
        if (!fp) { fprintf(stderr, "cannot open %s\n", argv[f]); return 2; }
        std::string seed;
        char buf[65536];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, fp)) > 0) seed.append(buf, n);
        fclose(fp);
        {   // the seed itself must parse
            ipcfp_parsed_bundle* pb = nullptr;
            if (ipcfp_bundle_from_json(seed.data(), seed.size(), &pb) != IPCFP_OK || !pb || check(pb)) { fprintf(stderr, "seed %s refused\n", argv[f]); return 1; }
            ipcfp_parsed_bundle_free(pb);
        }
        for (uint64_t m = 0; m < muts; m++) {
            std::string s = seed;
            const unsigned edits = 1 + (unsigned)(rnd() % 4);
            for (unsigned k = 0; k < edits && !s.empty(); k++) {
                const size_t pos = (size_t)(rnd() % s.size());
                switch (rnd() % 7) {
                    case 0: s[pos] = (char)rnd(); break;
                    case 1: s[pos] ^= (char)(1u << (rnd() % 8)); break;
                    case 2: { const char* d = DICT[rnd() % (sizeof DICT / sizeof *DICT)]; s.insert(pos, d); break; }
                    case 3: { const char* d = DICT[rnd() % (sizeof DICT / sizeof *DICT)]; size_t l = strlen(d); s.replace(pos, l < s.size() - pos ? l : s.size() - pos, d); break; }
                    case 4: { size_t l = 1 + (size_t)(rnd() % 64); s.erase(pos, l); break; }
                    case 5: { size_t l = 1 + (size_t)(rnd() % 256); if (l > s.size() - pos) l = s.size() - pos; s.insert(pos, s.substr(pos, l)); break; }
                    default: if (rnd() % 4 == 0) s.resize(pos); else s[pos] = "\"\\{}[],:0x"[rnd() % 10]; break;
                }
            }
            // an exact-size heap copy: one byte read past `len` is an ASan report
            char* text = (char*)malloc(s.size() ? s.size() : 1);
            memcpy(text, s.data(), s.size());
            ipcfp_parsed_bundle* pb = nullptr;
            ipcfp_status st = ipcfp_bundle_from_json(text, s.size(), &pb);
            total++;
            if (st == IPCFP_OK) {
                if (!pb) { fprintf(stderr, "OK without an object\n"); return 1; }
                if (check(pb)) { fprintf(stderr, "(mutation %llu of %s)\n", (unsigned long long)m, argv[f]); return 1; }
                ipcfp_parsed_bundle_free(pb);
                accepted++;
            } else {
                if (pb) { fprintf(stderr, "status %d with an object\n", (int)st); return 1; }
                if (st != IPCFP_ERR_INVALID_ARG && st != IPCFP_ERR_UNSUPPORTED) { fprintf(stderr, "undocumented status %d\n", (int)st); return 1; }
                refused++;
            }
            free(text);
        }
    }
    printf("ok: %llu mutated bundles parsed: %llu accepted (every promised byte in bounds), %llu refused with a documented status\n",
           (unsigned long long)total, (unsigned long long)accepted, (unsigned long long)refused);
    return 0;
}
