/* TEST INFRASTRUCTURE -- CPU restatement of the sketch construction behind Metalign's training database:
 *   CMash MakeStreamingDNADatabase.py <list> <out.h5> -n 1000 -k 60   (local_tests/retrain_and_test_metalign.sh:49)
 * i.e. MinHash.CountEstimator(n, max_prime=9999999999971, ksize, save_kmers='y', rev_comp=False).parse_file():
 *   for every record: seq.upper(), split on [^ACTG], every k-mer of every piece -> add(kmer):
 *       h = khmer.hash_no_rc_murmur3(kmer) % p           (MurmurHash3_x64_128, seed 0, first 64-bit word)
 *       if h >= mins[-1]: return
 *       i = bisect_left(mins, h); if mins[i] == h: counts[i] += 1
 *       else: insert (h, 1, kmer) at i, drop the last entry
 *   with mins = [p] * n, counts = [0] * n, kmers = [''] * n to start with.
 * CMash and khmer are third-party modules that are not in /root/reference and not installed here (SURVEY.md A.2,
 * [UPSTREAM]); the hash itself is pinned by MurmurHash3's published vectors (tests/test_sketch.py), the rest of the
 * restatement is PARITY UNPINNED.  Written byte-wise and streaming, unlike the GPU path (word-wise, threshold + sort).
 * Only tests/ and tests/bench_sketch.py's CPU leg may call this. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t fmix(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
static uint64_t le64(const unsigned char* p, int nbytes) {        /* little-endian load of nbytes <= 8 bytes */
    uint64_t v = 0;
    for (int i = nbytes - 1; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
/* MurmurHash3_x64_128 (Austin Appleby, public domain algorithm), both output words */
API void sko_murmur3_x64_128(const unsigned char* data, uint64_t len, uint32_t seed, uint64_t* out) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t nblocks = len / 16;
    for (uint64_t i = 0; i < nblocks; ++i) {
        uint64_t k1 = le64(data + 16 * i, 8), k2 = le64(data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const unsigned char* tail = data + 16 * nblocks;
    const int rem = (int)(len & 15);
    if (rem > 8) { uint64_t k2 = le64(tail + 8, rem - 8); k2 *= c2; k2 = rotl(k2, 33); k2 *= c1; h2 ^= k2; }
    if (rem > 0) { uint64_t k1 = le64(tail, rem > 8 ? 8 : rem); k1 *= c1; k1 = rotl(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix(h1); h2 = fmix(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

static int is_acgt(unsigned char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static unsigned char up(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }

/* one genome: text[0..len); mins/counts/kmers hold n slots (kmers: n*K bytes, NUL-filled when unused) */
API void sko_sketch_genome(const unsigned char* text, uint64_t len, uint32_t n, uint32_t K, uint64_t prime, uint64_t* mins,
                           uint32_t* counts, char* kmers) {
    for (uint32_t i = 0; i < n; ++i) { mins[i] = prime; counts[i] = 0; }
    memset(kmers, 0, (size_t)n * K);
    unsigned char* buf = (unsigned char*)malloc(K + 1);
    uint64_t run = 0;                                  /* length of the current run of ACGT characters */
    for (uint64_t pos = 0; pos < len; ++pos) {
        const unsigned char c = up(text[pos]);
        if (!is_acgt(c)) { run = 0; continue; }
        if (++run < K) continue;
        const uint64_t start = pos + 1 - K;
        for (uint32_t b = 0; b < K; ++b) buf[b] = up(text[start + b]);
        uint64_t out[2];
        sko_murmur3_x64_128(buf, K, 0, out);
        const uint64_t h = out[0] % prime;
        if (h >= mins[n - 1]) continue;
        uint32_t lo = 0, hi = n;                       /* bisect_left */
        while (lo < hi) { uint32_t mid = (lo + hi) / 2; if (mins[mid] < h) lo = mid + 1; else hi = mid; }
        if (mins[lo] == h) { counts[lo] += 1; continue; }
        memmove(mins + lo + 1, mins + lo, (size_t)(n - 1 - lo) * sizeof(uint64_t));
        memmove(counts + lo + 1, counts + lo, (size_t)(n - 1 - lo) * sizeof(uint32_t));
        memmove(kmers + (size_t)(lo + 1) * K, kmers + (size_t)lo * K, (size_t)(n - 1 - lo) * K);
        mins[lo] = h; counts[lo] = 1;
        memcpy(kmers + (size_t)lo * K, buf, K);
    }
    free(buf);
}
/* G genomes, one after the other (OpenMP over genomes) */
API void sko_sketch_genomes(const unsigned char* text, const uint64_t* off, uint32_t G, uint32_t n, uint32_t K, uint64_t prime,
                            uint64_t* mins, uint32_t* counts, char* kmers) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < (int64_t)G; ++g)
        sko_sketch_genome(text + off[g], off[g + 1] - off[g], n, K, prime, mins + (size_t)g * n, counts + (size_t)g * n,
                          kmers + (size_t)g * n * K);
}
