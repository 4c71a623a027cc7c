/* TEST INFRASTRUCTURE -- CPU oracle, C restatement (OpenMP).  PARITY UNPINNED.
 *
 * Restates the database-selection hot path of nlapier2/Metalign on the CPU:
 *   R1  scripts/select_db.py:50-52   kmc -k60 -ci2 -cs3   canonical K-mer counting over the reads
 *   R2  local_tests/dump_kmers.py:7-14 + local_tests/retrain_and_test_metalign.sh:66
 *                                    D = canonical form of every non-empty sketch slot
 *   R3  scripts/select_db.py:54-65   kmc_tools intersect + kmc_dump: I = {count >= ci_min} & D
 *   R4  scripts/select_db.py:73-76   StreamingQueryDNADatabase.py <I> <db> 30-60-10 -c 0 --sensitive:
 *                                    per record, per offset, smallest-k prefilter gate, then
 *                                    forward-first / reverse-complement-only-if-empty prefix matching
 *   R5  (same call)                  distinct hit prefixes / distinct sketch prefixes per genome, per k
 * following SURVEY.md section 3.3.  KMC and CMash themselves are third-party programs that are
 * neither vendored in the reference tree nor installed here, and the reference pins no version
 * and holds no golden vector for this path: "PARITY UNPINNED" -- this file is checked only
 * against the hand-derived cases in tests/golden/ and against oracle_py.py (an independent
 * set-based restatement of the same steps).
 *
 * It is deliberately built differently from the CUDA product path: unsigned __int128 rolling
 * k-mers, an open-addressing hash set for D, a plainly sorted array + binary search for the
 * prefix "trie", a literal sorted array for the zero-false-positive prefilter E0, and
 * sort+unique of (genome, k, prefix) records for the numerators.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  It is the checker and the timed CPU baseline, never the product path.
 *
 * Build: make -C oracle   ->  oracle/_build/liboracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

typedef unsigned __int128 u128;
#define ORC_API __attribute__((visibility("default")))
#define MAXK 8

typedef struct { u128 key; uint32_t g; uint32_t j; } pent;            /* stored-orientation sketch entry */
typedef struct { uint32_t g; uint32_t ki; u128 prefix; } hitrec;      /* one (genome, k, prefix) hit */

typedef struct orc_db {
    uint32_t G, n, K, nk; uint32_t ks[MAXK];
    u128 mask;
    uint64_t np; pent* P;            /* sorted by key */
    uint64_t nd; u128* D;            /* sorted distinct canonical keys */
    uint64_t hcap; uint32_t* hslot;  /* open addressing: index into D or EMPTY */
    uint64_t ne0; u128* E0;          /* sorted (with duplicates) k0-mers: S[:k0] and rc(S[:k0]) */
    int64_t* den_real;               /* G*nk distinct real prefixes */
    uint8_t* has_empty;              /* G */
} orc_db;

typedef struct orc_query {
    orc_db* db; int ci_min, gate, count_empty;
    uint32_t* cnt;                   /* nd occurrence counters */
    uint64_t n_kmers;
    int reduced;                     /* counters were replaced by an externally reduced table */
} orc_query;

#define EMPTY 0xFFFFFFFFu

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
static inline uint64_t hash128(u128 k) { return mix64((uint64_t)k ^ mix64((uint64_t)(k >> 64) + 0x632BE59BD9B4E019ull)); }

/* reverse complement of a k-base value (first base most significant) */
static inline u128 rc_val(u128 v, uint32_t k) {
    u128 r = 0;
    for (uint32_t i = 0; i < k; ++i) { r = (r << 2) | (3u - (uint32_t)(v & 3u)); v >>= 2; }
    return r;
}

/* ---------------------------------------------------------------- sorting */
static int cmp_pent(const void* a, const void* b) {
    const pent* x = (const pent*)a; const pent* y = (const pent*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    if (x->g != y->g) return x->g < y->g ? -1 : 1;
    return x->j < y->j ? -1 : (x->j > y->j);
}
static int cmp_u128(const void* a, const void* b) {
    u128 x = *(const u128*)a, y = *(const u128*)b; return x < y ? -1 : (x > y);
}
static int cmp_hit(const void* a, const void* b) {
    const hitrec* x = (const hitrec*)a; const hitrec* y = (const hitrec*)b;
    if (x->g != y->g) return x->g < y->g ? -1 : 1;
    if (x->ki != y->ki) return x->ki < y->ki ? -1 : 1;
    if (x->prefix != y->prefix) return x->prefix < y->prefix ? -1 : 1;
    return 0;
}

/* generic bucketed parallel sort: partition on the top 12 bits of a `bits`-wide key, qsort buckets */
#define NBKT 4096
static void psort(void* base, uint64_t n, size_t sz, int (*cmp)(const void*, const void*),
                  u128 (*keyof)(const void*), uint32_t bits) {
    if (n < (1u << 16) || bits < 12) { qsort(base, n, sz, cmp); return; }
    uint64_t* cnt = (uint64_t*)calloc(NBKT + 1, sizeof(uint64_t));
    char* src = (char*)base;
    uint32_t sh = bits - 12;
    for (uint64_t i = 0; i < n; ++i) cnt[(uint32_t)(keyof(src + i * sz) >> sh) + 1]++;
    for (uint32_t b = 0; b < NBKT; ++b) cnt[b + 1] += cnt[b];
    uint64_t* pos = (uint64_t*)malloc(NBKT * sizeof(uint64_t));
    memcpy(pos, cnt, NBKT * sizeof(uint64_t));
    char* tmp = (char*)malloc(n * sz);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t b = (uint32_t)(keyof(src + i * sz) >> sh);
        memcpy(tmp + pos[b]++ * sz, src + i * sz, sz);
    }
#pragma omp parallel for schedule(dynamic, 8)
    for (uint32_t b = 0; b < NBKT; ++b)
        qsort(tmp + cnt[b] * sz, cnt[b + 1] - cnt[b], sz, cmp);
    memcpy(base, tmp, n * sz);
    free(tmp); free(pos); free(cnt);
}
static u128 keyof_pent(const void* p) { return ((const pent*)p)->key; }
static u128 keyof_u128(const void* p) { return *(const u128*)p; }

/* first index i in P with P[i].key >= v */
static inline uint64_t lb_pent(const pent* P, uint64_t n, u128 v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (P[mid].key < v) lo = mid + 1; else hi = mid; }
    return lo;
}
static inline int in_sorted_u128(const u128* a, uint64_t n, u128 v) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo < n && a[lo] == v;
}

/* ---------------------------------------------------------------- database */
ORC_API orc_db* orc_db_build(const uint64_t* keys, uint32_t G, uint32_t n, uint32_t K,
                             const uint32_t* ks, uint32_t nk, int threads) {
    if (K < 1 || K > 63 || nk < 1 || nk > MAXK) return NULL;
    for (uint32_t i = 0; i < nk; ++i) if (ks[i] < 1 || ks[i] > K || (i && ks[i] <= ks[i - 1])) return NULL;
    if (threads > 0) omp_set_num_threads(threads);
    orc_db* db = (orc_db*)calloc(1, sizeof(orc_db));
    db->G = G; db->n = n; db->K = K; db->nk = nk; memcpy(db->ks, ks, nk * sizeof(uint32_t));
    db->mask = (((u128)1) << (2 * K)) - 1;
    uint64_t total = (uint64_t)G * n;
    /* P: non-empty slots in stored orientation */
    db->P = (pent*)malloc((total ? total : 1) * sizeof(pent));
    db->has_empty = (uint8_t*)calloc(G ? G : 1, 1);
    uint64_t np = 0;
    for (uint64_t s = 0; s < total; ++s) {
        if (keys[2 * s] == ~0ull) { db->has_empty[s / n] = 1; continue; }
        pent e; e.key = (((u128)keys[2 * s]) << 64) | keys[2 * s + 1]; e.g = (uint32_t)(s / n); e.j = (uint32_t)(s % n);
        db->P[np++] = e;
    }
    db->np = np;
    psort(db->P, np, sizeof(pent), cmp_pent, keyof_pent, 2 * K);
    /* D: distinct canonical keys, sorted */
    u128* C = (u128*)malloc((np ? np : 1) * sizeof(u128));
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < np; ++i) { u128 r = rc_val(db->P[i].key, K); C[i] = db->P[i].key < r ? db->P[i].key : r; }
    psort(C, np, sizeof(u128), cmp_u128, keyof_u128, 2 * K);
    uint64_t nd = 0;
    for (uint64_t i = 0; i < np; ++i) if (i == 0 || C[i] != C[i - 1]) C[nd++] = C[i];
    db->D = C; db->nd = nd;
    /* hash set over D */
    uint64_t cap = 16; while (cap < 2 * nd + 1) cap <<= 1;
    db->hcap = cap; db->hslot = (uint32_t*)malloc(cap * sizeof(uint32_t));
    memset(db->hslot, 0xFF, cap * sizeof(uint32_t));
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < nd; ++i) {
        uint64_t h = hash128(C[i]) & (cap - 1);
        for (;;) {
            uint32_t expect = EMPTY;
            if (__atomic_compare_exchange_n(&db->hslot[h], &expect, (uint32_t)i, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) break;
            h = (h + 1) & (cap - 1);
        }
    }
    /* E0: what a zero-false-positive prefilter holds for the smallest k */
    uint32_t k0 = ks[0];
    db->ne0 = 2 * np; db->E0 = (u128*)malloc((db->ne0 ? db->ne0 : 1) * sizeof(u128));
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < np; ++i) {
        u128 p0 = db->P[i].key >> (2 * (K - k0));
        db->E0[2 * i] = p0; db->E0[2 * i + 1] = rc_val(p0, k0);
    }
    psort(db->E0, db->ne0, sizeof(u128), cmp_u128, keyof_u128, 2 * k0);
    /* denominators: distinct real k-prefixes per genome */
    db->den_real = (int64_t*)calloc((size_t)(G ? G : 1) * nk, sizeof(int64_t));
    {   /* regroup P by genome */
        uint64_t* gstart = (uint64_t*)calloc((size_t)G + 1, sizeof(uint64_t));
        for (uint64_t i = 0; i < np; ++i) gstart[db->P[i].g + 1]++;
        for (uint32_t g = 0; g < G; ++g) gstart[g + 1] += gstart[g];
        uint64_t* fill = (uint64_t*)malloc(((size_t)G + 1) * sizeof(uint64_t));
        memcpy(fill, gstart, ((size_t)G + 1) * sizeof(uint64_t));
        u128* byg = (u128*)malloc((np ? np : 1) * sizeof(u128));
        for (uint64_t i = 0; i < np; ++i) byg[fill[db->P[i].g]++] = db->P[i].key;   /* stays key-sorted within g */
#pragma omp parallel for schedule(dynamic, 64)
        for (uint32_t g = 0; g < G; ++g) {
            for (uint32_t ki = 0; ki < nk; ++ki) {
                uint32_t sh = 2 * (K - ks[ki]); int64_t d = 0;
                for (uint64_t i = gstart[g]; i < gstart[g + 1]; ++i)
                    if (i == gstart[g] || (byg[i] >> sh) != (byg[i - 1] >> sh)) d++;
                db->den_real[(size_t)g * nk + ki] = d;
            }
        }
        free(byg); free(fill); free(gstart);
    }
    return db;
}

ORC_API void orc_db_free(orc_db* db) {
    if (!db) return;
    free(db->P); free(db->D); free(db->hslot); free(db->E0); free(db->den_real); free(db->has_empty); free(db);
}
ORC_API uint64_t orc_db_num_distinct(const orc_db* db) { return db->nd; }
ORC_API uint64_t orc_db_num_entries(const orc_db* db) { return db->np; }
/* sorted distinct canonical keys as (hi, lo) pairs */
ORC_API void orc_db_distinct_keys(const orc_db* db, uint64_t* out) {
    for (uint64_t i = 0; i < db->nd; ++i) { out[2 * i] = (uint64_t)(db->D[i] >> 64); out[2 * i + 1] = (uint64_t)db->D[i]; }
}

static inline int64_t d_lookup(const orc_db* db, u128 key) {
    uint64_t h = hash128(key) & (db->hcap - 1);
    for (;;) {
        uint32_t s = db->hslot[h];
        if (s == EMPTY) return -1;
        if (db->D[s] == key) return (int64_t)s;
        h = (h + 1) & (db->hcap - 1);
    }
}

/* ---------------------------------------------------------------- query: R1 */
ORC_API orc_query* orc_query_begin(orc_db* db, int ci_min, int gate_mode, int count_empty_in_den) {
    if (ci_min < 1 || (gate_mode != 0 && gate_mode != 1)) return NULL;
    orc_query* q = (orc_query*)calloc(1, sizeof(orc_query));
    q->db = db; q->ci_min = ci_min; q->gate = gate_mode; q->count_empty = count_empty_in_den;
    q->cnt = (uint32_t*)calloc(db->nd ? db->nd : 1, sizeof(uint32_t));
    return q;
}
ORC_API void orc_query_free(orc_query* q) { if (q) { free(q->cnt); free(q); } }

static inline void count_one(orc_query* q, u128 fwd, u128 rcv) {
    u128 c = fwd < rcv ? fwd : rcv;
    int64_t s = d_lookup(q->db, c);
    if (s >= 0) {
        uint32_t old = __atomic_load_n(&q->cnt[s], __ATOMIC_RELAXED);
        if (old < 0x7FFFFFFFu) __atomic_fetch_add(&q->cnt[s], 1u, __ATOMIC_RELAXED);
    }
}

/* reads as ASCII: read i = text[off[i] .. off[i+1]); anything outside ACGTacgt breaks the run */
ORC_API int orc_query_push_ascii(orc_query* q, const char* text, const uint64_t* off, uint64_t nreads) {
    const orc_db* db = q->db; const uint32_t K = db->K; const u128 mask = db->mask;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
    for (uint64_t r = 0; r < nreads; ++r) {
        u128 fwd = 0, rcv = 0; uint32_t run = 0;
        for (uint64_t p = off[r]; p < off[r + 1]; ++p) {
            uint32_t b;
            switch (text[p]) {
                case 'A': case 'a': b = 0; break;
                case 'C': case 'c': b = 1; break;
                case 'G': case 'g': b = 2; break;
                case 'T': case 't': b = 3; break;
                default: b = 4; break;
            }
            if (b == 4) { run = 0; fwd = 0; rcv = 0; continue; }
            fwd = ((fwd << 2) | b) & mask;
            rcv = (rcv >> 2) | (((u128)(3u - b)) << (2 * (K - 1)));
            if (++run >= K) { total++; count_one(q, fwd, rcv); }
        }
    }
    q->n_kmers += total;
    return 0;
}

/* reads 2-bit packed back to back (base i: byte i/4, bits 7-2(i%4)..6-2(i%4)); nmask bit i: byte i/8, bit 7-(i%8).
 * off (nreads+1, in bases) or NULL with fixed read_len. */
ORC_API int orc_query_push_packed(orc_query* q, const uint8_t* bases, const uint8_t* nmask,
                                  const uint64_t* off, uint64_t nreads, uint32_t read_len) {
    const orc_db* db = q->db; const uint32_t K = db->K; const u128 mask = db->mask;
    uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : total)
    for (uint64_t r = 0; r < nreads; ++r) {
        uint64_t p0 = off ? off[r] : r * (uint64_t)read_len, p1 = off ? off[r + 1] : (r + 1) * (uint64_t)read_len;
        u128 fwd = 0, rcv = 0; uint32_t run = 0;
        for (uint64_t p = p0; p < p1; ++p) {
            if (nmask && ((nmask[p >> 3] >> (7 - (p & 7))) & 1)) { run = 0; continue; }
            uint32_t b = (bases[p >> 2] >> (6 - 2 * (p & 3))) & 3u;
            fwd = ((fwd << 2) | b) & mask;
            rcv = (rcv >> 2) | (((u128)(3u - b)) << (2 * (K - 1)));
            if (++run >= K) { total++; count_one(q, fwd, rcv); }
        }
    }
    q->n_kmers += total;
    return 0;
}

/* multi-rank seam: per-rank counters clamped to ci_min, summed elsewhere, written back */
ORC_API void orc_query_export_counts(const orc_query* q, uint8_t* out) {
    for (uint64_t i = 0; i < q->db->nd; ++i) {
        uint32_t c = q->cnt[i]; out[i] = (uint8_t)(c > (uint32_t)q->ci_min ? (uint32_t)q->ci_min : c);
    }
}
ORC_API void orc_query_import_counts(orc_query* q, const uint8_t* in) {
    for (uint64_t i = 0; i < q->db->nd; ++i) q->cnt[i] = in[i];
    q->reduced = 1;
}

/* ---------------------------------------------------------------- query: R3-R5 */
typedef struct { hitrec* v; uint64_t n, cap; } hitvec;
static inline void hv_push(hitvec* h, uint32_t g, uint32_t ki, u128 prefix) {
    if (h->n == h->cap) { h->cap = h->cap ? 2 * h->cap : 1024; h->v = (hitrec*)realloc(h->v, h->cap * sizeof(hitrec)); }
    h->v[h->n].g = g; h->v[h->n].ki = ki; h->v[h->n].prefix = prefix; h->n++;
}

/* match(w): forward prefix range first; reverse complement only if forward is empty */
static inline int match_push(const orc_db* db, hitvec* hv, u128 w, uint32_t ki) {
    uint32_t k = db->ks[ki]; uint32_t sh = 2 * (db->K - k);
    for (int pass = 0; pass < 2; ++pass) {
        u128 v = pass == 0 ? w : rc_val(w, k);
        u128 lo = v << sh, hi = lo + (((u128)1) << sh);
        uint64_t i = lb_pent(db->P, db->np, lo); int any = 0;
        for (; i < db->np && db->P[i].key < hi; ++i) { hv_push(hv, db->P[i].g, ki, db->P[i].key >> sh); any = 1; }
        if (any) return 1;
    }
    return 0;
}

ORC_API int orc_query_finish(orc_query* q, int64_t* num, int64_t* den, double* ci,
                             uint64_t* n_intersect, uint64_t* n_kmers) {
    const orc_db* db = q->db; const uint32_t K = db->K, nk = db->nk, k0 = db->ks[0];
    int nth = omp_get_max_threads();
    hitvec* hv = (hitvec*)calloc(nth, sizeof(hitvec));
    uint64_t ni = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : ni)
    for (uint64_t d = 0; d < db->nd; ++d) {
        if (q->cnt[d] < (uint32_t)q->ci_min) continue;
        ni++;
        hitvec* h = &hv[omp_get_thread_num()];
        u128 x = db->D[d];
        for (uint32_t i = 0; i + k0 <= K; ++i) {
            u128 w0 = (x >> (2 * (K - i - k0))) & ((((u128)1) << (2 * k0)) - 1);
            int possible = q->gate == 1 ? 1 : in_sorted_u128(db->E0, db->ne0, w0);
            if (!possible) continue;
            match_push(db, h, w0, 0);
            for (uint32_t ki = 1; ki < nk; ++ki) {
                uint32_t k = db->ks[ki];
                if (i + k > K) continue;
                u128 wk = (x >> (2 * (K - i - k))) & ((((u128)1) << (2 * k)) - 1);
                match_push(db, h, wk, ki);
            }
        }
    }
    uint64_t tot = 0; for (int t = 0; t < nth; ++t) tot += hv[t].n;
    hitrec* all = (hitrec*)malloc((tot ? tot : 1) * sizeof(hitrec));
    uint64_t o = 0; for (int t = 0; t < nth; ++t) { if (hv[t].n) memcpy(all + o, hv[t].v, hv[t].n * sizeof(hitrec)); o += hv[t].n; free(hv[t].v); }
    free(hv);
    qsort(all, tot, sizeof(hitrec), cmp_hit);
    size_t cells = (size_t)db->G * nk;
    memset(num, 0, cells * sizeof(int64_t));
    for (uint64_t i = 0; i < tot; ++i)
        if (i == 0 || cmp_hit(&all[i], &all[i - 1]) != 0) num[(size_t)all[i].g * nk + all[i].ki]++;
    free(all);
    for (uint32_t g = 0; g < db->G; ++g)
        for (uint32_t ki = 0; ki < nk; ++ki) {
            size_t c = (size_t)g * nk + ki;
            den[c] = db->den_real[c] + ((q->count_empty && db->has_empty[g]) ? 1 : 0);
            ci[c] = num[c] > 0 ? (double)num[c] / (double)den[c] : 0.0;
        }
    if (n_intersect) *n_intersect = ni;
    if (n_kmers) *n_kmers = q->n_kmers;
    return 0;
}

/* I as sorted (hi, lo) canonical keys; returns |I| (writes at most cap pairs) */
ORC_API uint64_t orc_query_intersection(const orc_query* q, uint64_t* out, uint64_t cap) {
    uint64_t ni = 0;
    for (uint64_t d = 0; d < q->db->nd; ++d) {
        if (q->cnt[d] < (uint32_t)q->ci_min) continue;
        if (ni < cap) { out[2 * ni] = (uint64_t)(q->db->D[d] >> 64); out[2 * ni + 1] = (uint64_t)q->db->D[d]; }
        ni++;
    }
    return ni;
}

ORC_API int orc_max_threads(void) { return omp_get_max_threads(); }
ORC_API void orc_set_threads(int t) { if (t > 0) omp_set_num_threads(t); }
