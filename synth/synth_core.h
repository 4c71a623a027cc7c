/* Synthetic workload generator -- shared core.
 *
 * TEST / BENCH INFRASTRUCTURE, not the product path and not the oracle: it only
 * manufactures inputs (a sketch database and reads simulated from its genomes).
 * The same source is compiled for the host (synth_cpu.c) and for the device
 * (synth_cuda.cu), so both produce bit-identical data from the same parameters.
 *
 * Everything is a pure function of (seed, index): genomes are never materialised.
 *   - genome g belongs to a block of `strain_period` genomes; the last genome of a
 *     block is a "strain" of the first (its root) with ~0.1 % substitutions, all
 *     others are their own root.  strain_period == 0 disables strains.
 *   - root genome length is uniform in [len_min, len_max]; base i of root r is two
 *     bits of a counter-based hash of (r, i/32).
 *   - sketch slot j of genome g is the K-mer at position hash(root(g), j) -- the SAME
 *     position for a strain and its root, so related genomes share most sketch k-mers
 *     (as real bottom-n MinHash sketches of related genomes do).  Slots are stored in
 *     genome-strand orientation.  `tiny_pct` percent of genomes keep only a fraction of
 *     their slots; the rest are empty ('' in CMash) and are emitted as hi = lo = ~0.
 *   - reads: genome drawn from `n_present` present genomes with a skewed integer
 *     abundance, uniform start, random strand, fixed length L, 0.5 % substitutions,
 *     0.1 % N.  paired != 0: reads 2f and 2f+1 are the two ends of fragment f.
 */
#ifndef MLG_SYNTH_CORE_H
#define MLG_SYNTH_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD static inline
#endif

typedef struct {
    uint64_t seed;
    uint32_t G;             /* genomes */
    uint32_t n;             /* sketch slots per genome */
    uint32_t K;             /* sketch k-mer length (<= 63) */
    uint32_t strain_period; /* 0 = no strains; else genome g with g % period == period-1 is a strain of g-(period-1) */
    uint32_t tiny_pct;      /* percent of genomes with a partly empty sketch */
    uint32_t len_min;       /* root genome length range, bases */
    uint32_t len_max;
    uint32_t n_present;     /* genomes the reads are drawn from */
    uint32_t read_len;      /* L */
    uint32_t paired;        /* 0/1 */
    uint32_t sub_per_64k;   /* substitution rate, in 1/65536 per base (328 ~ 0.5 %) */
    uint32_t n_per_64k;     /* N rate, in 1/65536 per base (66 ~ 0.1 %) */
} syn_params;

SYN_HD uint64_t syn_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

SYN_HD uint64_t syn_rnd(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b) {
    return syn_mix(syn_mix(syn_mix(seed ^ (stream * 0xD6E8FEB86659FD93ull)) + a) + b);
}

SYN_HD int syn_is_strain(const syn_params* p, uint32_t g) {
    return p->strain_period >= 2 && (g % p->strain_period) == p->strain_period - 1;
}
SYN_HD uint32_t syn_root(const syn_params* p, uint32_t g) {
    return syn_is_strain(p, g) ? g - (p->strain_period - 1) : g;
}
SYN_HD uint32_t syn_genome_len(const syn_params* p, uint32_t root) {
    uint32_t span = p->len_max - p->len_min + 1;
    return p->len_min + (uint32_t)(syn_rnd(p->seed, 3, root, 0) % span);
}
/* base (0..3) of ROOT genome r at position pos */
SYN_HD uint32_t syn_root_base(const syn_params* p, uint32_t r, uint32_t pos) {
    uint64_t blk = syn_rnd(p->seed, 4, r, pos >> 5);
    return (uint32_t)(blk >> (2 * (pos & 31))) & 3u;
}
/* base of genome g (root or strain) at position pos */
SYN_HD uint32_t syn_genome_base(const syn_params* p, uint32_t g, uint32_t pos) {
    uint32_t r = syn_root(p, g);
    uint32_t b = syn_root_base(p, r, pos);
    if (r != g) {
        uint64_t m = syn_rnd(p->seed, 5, g, pos >> 2);
        uint32_t f = (uint32_t)(m >> (16 * (pos & 3))) & 0xFFFFu;
        if (f < 66u) b = (b + 1u + f % 3u) & 3u;
    }
    return b;
}
/* same function, with the two hash words cached across neighbouring positions */
typedef struct { uint32_t g, r; uint32_t blk_idx, mut_idx; uint64_t blk, mut; } syn_gcache;
SYN_HD void syn_gcache_init(const syn_params* p, syn_gcache* c, uint32_t g) {
    c->g = g; c->r = syn_root(p, g); c->blk_idx = 0xFFFFFFFFu; c->mut_idx = 0xFFFFFFFFu; c->blk = 0; c->mut = 0;
}
SYN_HD uint32_t syn_genome_base_c(const syn_params* p, syn_gcache* c, uint32_t pos) {
    if ((pos >> 5) != c->blk_idx) { c->blk_idx = pos >> 5; c->blk = syn_rnd(p->seed, 4, c->r, pos >> 5); }
    uint32_t b = (uint32_t)(c->blk >> (2 * (pos & 31))) & 3u;
    if (c->r != c->g) {
        if ((pos >> 2) != c->mut_idx) { c->mut_idx = pos >> 2; c->mut = syn_rnd(p->seed, 5, c->g, pos >> 2); }
        uint32_t f = (uint32_t)(c->mut >> (16 * (pos & 3))) & 0xFFFFu;
        if (f < 66u) b = (b + 1u + f % 3u) & 3u;
    }
    return b;
}
/* number of real (non-empty) sketch slots of genome g */
SYN_HD uint32_t syn_real_slots(const syn_params* p, uint32_t g) {
    if (p->tiny_pct && (syn_rnd(p->seed, 7, g, 0) % 100u) < p->tiny_pct) {
        uint32_t lo = p->n / 5u;
        uint32_t span = (p->n * 7u) / 10u + 1u;
        return lo + (uint32_t)(syn_rnd(p->seed, 8, g, 0) % span);
    }
    return p->n;
}
/* sketch slot (g, j) -> 2K-bit key, first base most significant; empty -> all ones */
SYN_HD void syn_sketch_key(const syn_params* p, uint32_t g, uint32_t j, uint64_t* hi, uint64_t* lo) {
    if (j >= syn_real_slots(p, g)) { *hi = ~0ull; *lo = ~0ull; return; }
    uint32_t r = syn_root(p, g);
    uint32_t len = syn_genome_len(p, r);
    uint32_t pos = (uint32_t)(syn_rnd(p->seed, 6, r, j) % (uint64_t)(len - p->K + 1u));
    uint64_t h = 0, l = 0;
    syn_gcache gc; syn_gcache_init(p, &gc, g);
    for (uint32_t t = 0; t < p->K; ++t) {
        uint32_t b = syn_genome_base_c(p, &gc, pos + t);
        h = (h << 2) | (l >> 62);
        l = (l << 2) | b;
    }
    *hi = h; *lo = l;
}

/* ---- reads ---------------------------------------------------------------- */
SYN_HD uint32_t syn_present_genome(const syn_params* p, uint32_t i) {
    return (uint32_t)(syn_rnd(p->seed, 9, i, 0) % p->G);
}
SYN_HD uint64_t syn_present_weight(const syn_params* p, uint32_t i) {
    uint64_t u = syn_rnd(p->seed, 10, i, 0) % 1000u;
    return 1u + (u * u) / 1000u;
}

typedef struct { uint32_t g; uint32_t start; uint32_t rev; } syn_read_src;

/* cum[i] = sum of weights of present genomes 0..i (inclusive), length n_present */
SYN_HD syn_read_src syn_read_source(const syn_params* p, const uint64_t* cum, uint64_t r) {
    syn_read_src s;
    uint64_t total = cum[p->n_present - 1];
    uint64_t unit = p->paired ? (r >> 1) : r;
    uint64_t u = syn_rnd(p->seed, 11, unit, 0) % total;
    uint32_t lo = 0, hi = p->n_present - 1;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cum[mid] > u) hi = mid; else lo = mid + 1; }
    s.g = syn_present_genome(p, lo);
    uint32_t len = syn_genome_len(p, syn_root(p, s.g));
    uint32_t L = p->read_len;
    if (!p->paired) {
        s.start = (uint32_t)(syn_rnd(p->seed, 12, r, 0) % (uint64_t)(len - L + 1u));
        s.rev = (uint32_t)(syn_rnd(p->seed, 13, r, 0) & 1u);
    } else {
        uint32_t flen = 2u * L - 30u + (uint32_t)(syn_rnd(p->seed, 15, unit, 0) % 61u);
        if (flen > len) flen = len;
        if (flen < L) flen = L;
        uint32_t fstart = (uint32_t)(syn_rnd(p->seed, 12, unit, 0) % (uint64_t)(len - flen + 1u));
        uint32_t fstrand = (uint32_t)(syn_rnd(p->seed, 13, unit, 0) & 1u);
        uint32_t mate = (uint32_t)(r & 1u);
        if ((mate ^ fstrand) == 0) { s.start = fstart; s.rev = 0; }
        else { s.start = fstart + flen - L; s.rev = 1; }
    }
    return s;
}
/* base t (0..L-1) of read r: returns 0..3, or 4 for N.  gc must have been initialised for s->g. */
SYN_HD uint32_t syn_read_base(const syn_params* p, const syn_read_src* s, syn_gcache* gc, uint64_t r, uint32_t t) {
    uint32_t L = p->read_len;
    uint32_t gpos = s->rev ? (s->start + (L - 1u - t)) : (s->start + t);
    uint32_t b = syn_genome_base_c(p, gc, gpos);
    if (s->rev) b = 3u - b;
    uint64_t e = syn_rnd(p->seed, 14, r, t >> 2);
    uint32_t f = (uint32_t)(e >> (16 * (t & 3))) & 0xFFFFu;
    if (f < p->sub_per_64k) b = (b + 1u + f % 3u) & 3u;
    else if (f < p->sub_per_64k + p->n_per_64k) b = 4u;
    return b;
}

#endif /* MLG_SYNTH_CORE_H */
