// Reader of KMC k-mer databases (<prefix>.kmc_pre + <prefix>.kmc_suf), both layouts: version 0 ("KMC1", what kmc_tools
// writes) and 0x200 ("KMC2", what kmc writes).  These are the artefacts the reference moves between its subprocesses:
// data/cmash_db_n1000_k60_dump (scripts/select_db.py:44; made at local_tests/retrain_and_test_metalign.sh:66),
// reads_60mers and 60mers_intersection (select_db.py:50-56).  The product does not need them to run -- it is used to
// check a native database against the KMC database shipped beside it (scripts/check_db_against_kmc.py) and to read a
// real reference run's intersection for comparison.
// KMC is not vendored under /root/reference and not installed here: the layout follows the KMC API documentation
// ("k-mer database format") as recalled in SURVEY.md A.1 and is UNPINNED against files written by a real kmc; the
// tests use an independent pure-Python encoder (tests/kmcdb.py).
//
//   .kmc_pre  "KMCP" | u64 LUT[] (+ guard) | [u32 signature map, 0x200 only] | header | u32 header_bytes | "KMCP"
//             header: u32 k, mode, counter_size, lut_prefix_length, [signature_len], min_count, max_count; u64 total;
//                     u8 !both_strands; 3 pad; u32 max_count_hi; 20 reserved; u32 kmc_version
//             LUT entry i = records of .kmc_suf before those of prefix (i mod 4^lut_prefix_length)
//   .kmc_suf  "KMCS" | records: (k - lut_prefix_length) / 4 suffix bytes, first base in the top bits; counter_size bytes LE | "KMCS"
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

namespace kmcdb {

struct Info {
    uint32_t k = 0, mode = 0, counter_size = 0, lut_prefix_length = 0, signature_len = 0, min_count = 0, version = 0;
    uint64_t max_count = 0, total = 0;
    bool canonical = true;
};

struct Reader {
    Info info;
    std::vector<unsigned char> pre, suf;
    std::string error;
    size_t lut_entries = 0;

    static bool slurp(const std::string& path, std::vector<unsigned char>& out) {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END);
        const long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        out.resize(n > 0 ? (size_t)n : 0);
        const bool ok = n >= 0 && fread(out.data(), 1, out.size(), f) == out.size();
        fclose(f);
        return ok;
    }
    template <typename T>
    T at(const std::vector<unsigned char>& v, size_t off) const { T x; memcpy(&x, v.data() + off, sizeof(T)); return x; }

    bool open(const std::string& prefix) {
        if (!slurp(prefix + ".kmc_pre", pre) || !slurp(prefix + ".kmc_suf", suf)) { error = "cannot read " + prefix + ".kmc_pre / .kmc_suf"; return false; }
        if (pre.size() < 12 + 44 || memcmp(pre.data(), "KMCP", 4) || memcmp(pre.data() + pre.size() - 4, "KMCP", 4) ||
            suf.size() < 8 || memcmp(suf.data(), "KMCS", 4) || memcmp(suf.data() + suf.size() - 4, "KMCS", 4)) { error = prefix + ": KMC markers missing"; return false; }
        info.version = at<uint32_t>(pre, pre.size() - 12);
        const uint32_t hbytes = at<uint32_t>(pre, pre.size() - 8);
        if (info.version != 0 && info.version != 0x200) { error = prefix + ": unknown kmc_version"; return false; }
        if ((size_t)hbytes + 12 > pre.size() || hbytes < 44) { error = prefix + ": bad header offset"; return false; }
        size_t o = pre.size() - 8 - hbytes;
        const size_t h0 = o;
        info.k = at<uint32_t>(pre, o); info.mode = at<uint32_t>(pre, o + 4); info.counter_size = at<uint32_t>(pre, o + 8);
        info.lut_prefix_length = at<uint32_t>(pre, o + 12);
        o += 16;
        if (info.version == 0x200) { info.signature_len = at<uint32_t>(pre, o); o += 4; }
        info.min_count = at<uint32_t>(pre, o); info.max_count = at<uint32_t>(pre, o + 4); info.total = at<uint64_t>(pre, o + 8);
        info.canonical = pre[o + 16] == 0;
        info.max_count |= (uint64_t)at<uint32_t>(pre, o + 20) << 32;
        if (info.k == 0 || info.k > 63 || info.lut_prefix_length == 0 || info.lut_prefix_length >= info.k || info.lut_prefix_length > 15 ||
            (info.k - info.lut_prefix_length) % 4 || info.counter_size > 8 || info.mode != 0 || info.signature_len > 11) {
            error = prefix + ": unsupported parameters (k <= 63, k-mer counters only)"; return false;
        }
        size_t lut_end = h0;
        if (info.version == 0x200) {
            const size_t map_bytes = ((size_t)1 << (2 * info.signature_len)) * 4 + 4;
            if (map_bytes + 4 > lut_end) { error = prefix + ": signature map larger than the file"; return false; }
            lut_end -= map_bytes;
        }
        lut_entries = (lut_end - 4) / 8;
        if (lut_entries < 2) { error = prefix + ": empty prefix table"; return false; }
        const size_t rec = (info.k - info.lut_prefix_length) / 4 + info.counter_size;
        if (info.total > suf.size() || 8 + info.total * rec != suf.size()) { error = prefix + ": .kmc_suf size does not match total_kmers"; return false; }
        return true;
    }
    // every k-mer as a 2k-bit integer (hi, lo; first base most significant), in file order; counts may be null
    bool read_all(uint64_t* keys, uint32_t* counts) {
        const uint32_t p = info.lut_prefix_length, sb = (info.k - p) / 4, cs = info.counter_size;
        const size_t rec = sb + cs, per = (size_t)1 << (2 * p);
        const unsigned suf_bits = 2 * (info.k - p);
        uint64_t done = 0;
        for (size_t i = 0; i + 1 < lut_entries; ++i) {
            const uint64_t a = at<uint64_t>(pre, 4 + 8 * i), b = at<uint64_t>(pre, 4 + 8 * (i + 1));
            if (b < a || b > info.total) { error = "prefix table is not monotone"; return false; }
            const unsigned __int128 pv = (unsigned __int128)(i % per) << suf_bits;
            for (uint64_t r = a; r < b; ++r) {
                const unsigned char* q = suf.data() + 4 + r * rec;
                unsigned __int128 v = 0;
                for (uint32_t j = 0; j < sb; ++j) v = (v << 8) | q[j];
                v |= pv;
                keys[2 * r] = (uint64_t)(v >> 64); keys[2 * r + 1] = (uint64_t)v;
                if (counts) { uint64_t c = 0; for (uint32_t j = 0; j < cs; ++j) c |= (uint64_t)q[sb + j] << (8 * j); counts[r] = c > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c; }
                ++done;
            }
        }
        if (done != info.total) { error = "prefix tables do not cover total_kmers"; return false; }
        return true;
    }
};

}  // namespace kmcdb
