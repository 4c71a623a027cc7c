// Database build: turns G*n sketch slots (stored-orientation K-mers, '' = empty) into the
// HBM-resident structures the query kernels use.  Replaces, for this path, CMash's training HDF5 +
// marisa-trie `.tst` + hydra Bloom prefilter (scripts/select_db.py:69-70) and the KMC database of the
// dumped sketch k-mers (scripts/select_db.py:44; local_tests/retrain_and_test_metalign.sh:49-66).
//
//   P   all non-empty slots sorted by key + slot id payload, with a bucket index on the top bits of
//       the key ("bucketed binary search": a k-prefix query is a key range, every k served by one array)
//   rep per queried k: representative slot of each (genome, k-prefix) class -> numerators dedupe by
//       prefix exactly like CMash's per-genome `unique_kmers` set; class counts are the denominators
//   D   distinct canonical keys, sorted by a 64-bit hash; level-1 table of 16/32-byte buckets of 31-bit
//       fingerprints so that a read k-mer that is NOT in the database (>99.9 % of them) costs exactly
//       one DRAM sector; bstart maps a bucket back to its run of D for exact verification
//
// Everything runs on the device (CUB radix sorts / scans + small kernels).
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "mlg_internal.h"

namespace {

constexpr int TPB = 256;
inline unsigned nblk(unsigned long long n) { return (unsigned)((n + TPB - 1) / TPB); }

// ---------------------------------------------------------------- small kernels
__global__ void k_mark_nonempty(const key128* keys, unsigned long long total, uint32_t n, unsigned char* flag,
                                unsigned char* has_empty) {
    unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (s >= total) return;
    bool e = key_is_empty(keys[s]);
    flag[s] = e ? 0 : 1;
    if (e) has_empty[s / n] = 1;
}
__global__ void k_gather_split(const key128* keys, const uint32_t* slots, uint32_t np, unsigned long long* hi,
                               unsigned long long* lo) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    key128 k = keys[slots[i]];
    hi[i] = k.hi; lo[i] = k.lo;
}
__global__ void k_iota(uint32_t* a, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
__global__ void k_gather_u64(const unsigned long long* src, const uint32_t* idx, uint32_t n, unsigned long long* dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t n, uint32_t* dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
// out[i] = {hi[i], lo_sorted[perm[i]]}
__global__ void k_zip_keys(const unsigned long long* hi, const unsigned long long* lo_sorted, const uint32_t* perm,
                           uint32_t n, key128* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key128 k; k.hi = hi[i]; k.lo = lo_sorted[perm ? perm[i] : i];
    out[i] = k;
}
__global__ void k_pbucket_hist(const key128* P, uint32_t np, uint32_t K, uint32_t pbits, uint32_t* counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    key128 b = key_shr(P[i], 2 * K - pbits);
    atomicAdd(&counts[(uint32_t)b.lo], 1u);
}
__global__ void k_slot_to_genome(const uint32_t* slot, uint32_t np, uint32_t n, uint32_t* g) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) g[i] = slot[i] / n;
}
// q-order = entries sorted by (genome, key).  A (genome, k-prefix) class is a run in q-order; its
// representative is the slot of the run head.
__global__ void k_rep_classes(const key128* P, const uint32_t* P_slot, const uint32_t* q2p, uint32_t np, uint32_t n,
                              uint32_t K, uint32_t k, uint32_t ki, uint32_t nk, uint32_t* rep_k, long long* den_real) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= np) return;
    uint32_t i = q2p[q];
    uint32_t slot = P_slot[i];
    uint32_t g = slot / n;
    key128 pre = key_prefix(P[i], K, k);
    uint32_t h = q;
    while (h > 0) {
        uint32_t ip = q2p[h - 1];
        if (P_slot[ip] / n != g) break;
        if (!key_eq(key_prefix(P[ip], K, k), pre)) break;
        --h;
    }
    rep_k[slot] = P_slot[q2p[h]];
    if (h == q) atomicAdd((unsigned long long*)&den_real[(size_t)g * nk + ki], 1ull);
}
__global__ void k_canon_split(const key128* P, uint32_t np, uint32_t K, unsigned long long* hi, unsigned long long* lo) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    key128 c = key_canon(P[i], K);
    hi[i] = c.hi; lo[i] = c.lo;
}
__global__ void k_flag_heads(const key128* a, uint32_t n, unsigned char* flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (i == 0 || !key_eq(a[i], a[i - 1])) ? 1 : 0;
}
// length of the run of equal keys that starts at every head, saturating at 255 (the number of sketch slots that hold the K-mer)
__global__ void k_run_lengths(const unsigned char* head, uint32_t n, unsigned char* len) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t l = 0;
    if (head[i]) { l = 1; while (l < 255u && i + l < n && !head[i + l]) ++l; }
    len[i] = (unsigned char)l;
}
__global__ void k_gather_u8(const unsigned char* src, const uint32_t* idx, uint32_t n, unsigned char* dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_hash_keys(const key128* a, uint32_t n, uint32_t K, uint32_t layout, uint32_t bbits, unsigned long long* h) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h[i] = layout == 2 ? key_hash_mz(a[i], K) : layout == 1 ? key_hash_sk(a[i], K, bbits) : key_hash(a[i], K);
}
// layout 2: level-1 bits of the identities of every K-mer's leftmost and rightmost minimum (kmer.cuh), and the alias
// entries of the K-mers whose two identities differ (pass 0 counts them, pass 1 writes them)
__global__ void k_fill_mzbits(const key128* D, uint32_t nd, uint32_t K, uint32_t fbits, uint32_t* MB, int pass,
                              unsigned long long* alias_z, uint32_t* alias_i, uint32_t* alias_bloom, unsigned long long* counter) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    unsigned long long zL, zR;
    key_mz(D[i], K, &zL, &zR);
    if (pass == 0) {
        const unsigned long long bl = mz_bit_index(zL, fbits), br = mz_bit_index(zR, fbits);
        atomicOr(&MB[bl >> 5], (1u << (bl & 31ull)) | (1u << mz_bit2((uint32_t)(zL >> 32))));
        if (zR != zL) { atomicOr(&MB[br >> 5], (1u << (br & 31ull)) | (1u << mz_bit2((uint32_t)(zR >> 32)))); atomicAdd(counter, 1ull); }
    } else if (zR != zL) {
        const unsigned long long j = atomicAdd(counter, 1ull);
        alias_z[j] = zR; alias_i[j] = i;
        const uint32_t zlo = (uint32_t)zR;
        atomicOr(&alias_bloom[(zlo & 0xFFFFu) >> 5], 1u << (zlo & 31u));
    }
}
__global__ void k_gather_key(const key128* src, const uint32_t* idx, uint32_t n, key128* dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
__global__ void k_hbucket_hist(const unsigned long long* h, uint32_t nd, uint32_t bbits, uint32_t* counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nd) atomicAdd(&counts[hash_bucket(h[i], bbits)], 1u);
}
__global__ void k_fill_filter(const unsigned long long* h, uint32_t nd, uint32_t nfw, uint32_t fk, uint32_t* F) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nd) atomicOr(&F[filter_word(h[i], nfw)], filter_mask(h[i], fk));
}
__global__ void k_fill_t1(const unsigned long long* h, uint32_t nd, uint32_t bbits, const uint32_t* bstart,
                          uint32_t slots, uint32_t layout, uint32_t* T1) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nd) return;
    unsigned long long b = hash_bucket(h[e], bbits);
    uint32_t s = e - bstart[b];
    uint32_t c = bstart[b + 1] - bstart[b];
    if (s < slots) {
        uint32_t v = hash_fp(h[e]);
        // more entries than slots: the probe must take the exact path.  Layout 0 flags it in bit 31 of the first slot;
        // layout 1 needs no flag, it treats every FULL bucket as overflowed
        if (s == 0 && c > slots && layout == 0) v |= 0x80000000u;
        T1[b * slots + s] = v;
    }
}

struct IsNonEmpty {
    const unsigned char* flag;
    __host__ __device__ bool operator()(const uint32_t& s) const { return flag[s] != 0; }
};

// exclusive prefix sum of counts[0..m) into out[0..m], out[m] = total
int exclusive_scan_u32(uint32_t* counts_inout_m_plus_1, size_t m, cudaStream_t st) {
    // counts has m+1 entries with counts[m] == 0; in-place exclusive sum over m+1 entries gives the bucket starts
    void* tmp = nullptr; size_t tb = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts_inout_m_plus_1, counts_inout_m_plus_1, m + 1, st));
    CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, counts_inout_m_plus_1, counts_inout_m_plus_1, m + 1, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) { mlg_set_error("DeviceScan failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
    return MLG_OK;
}

// stable radix sort of (u64 key, u32 value) pairs; results land in keys_out / vals_out
int sort_pairs_u64(const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                   uint32_t* vals_out, uint32_t n, int begin_bit, int end_bit, cudaStream_t st) {
    void* tmp = nullptr; size_t tb = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, st));
    CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, vals_in, vals_out, (int)n, begin_bit, end_bit, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) { mlg_set_error("DeviceRadixSort failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
    return MLG_OK;
}
int sort_pairs_u32(const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out, uint32_t n,
                   int end_bit, cudaStream_t st) {
    void* tmp = nullptr; size_t tb = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st));
    CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys_in, keys_out, vals_in, vals_out, (int)n, 0, end_bit, st);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    if (e != cudaSuccess) { mlg_set_error("DeviceRadixSort failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
    return MLG_OK;
}

// Sort n 2K-bit keys given as split (hi, lo) arrays; on return out[i] is the i-th smallest key and, if
// vals is non-null, vals_out[i] the value that travelled with it.  LSD: by lo, then stably by hi.
int sort_keys128(unsigned long long* hi, unsigned long long* lo, const uint32_t* vals, uint32_t n, uint32_t K,
                 key128* out, uint32_t* vals_out, cudaStream_t st) {
    if (n == 0) return MLG_OK;
    DevBuf<unsigned long long> lo_s, hi_g, hi_s;
    DevBuf<uint32_t> iota, perm1, perm2;
    MLG_TRY(lo_s.alloc(n)); MLG_TRY(iota.alloc(n)); MLG_TRY(perm1.alloc(n));
    k_iota<<<nblk(n), TPB, 0, st>>>(iota.p, n);
    int lo_bits = (2 * K < 64) ? (int)(2 * K) : 64;
    MLG_TRY(sort_pairs_u64(lo, lo_s.p, iota.p, perm1.p, n, 0, lo_bits, st));
    const uint32_t* final_perm = perm1.p;   // position in sorted order -> original index
    if (2 * K > 64) {
        MLG_TRY(hi_g.alloc(n)); MLG_TRY(hi_s.alloc(n)); MLG_TRY(perm2.alloc(n));
        k_gather_u64<<<nblk(n), TPB, 0, st>>>(hi, perm1.p, n, hi_g.p);
        MLG_TRY(sort_pairs_u64(hi_g.p, hi_s.p, iota.p, perm2.p, n, 0, (int)(2 * K - 64), st));
        // perm2[i] = position in the lo-sorted order
        k_zip_keys<<<nblk(n), TPB, 0, st>>>(hi_s.p, lo_s.p, perm2.p, n, out);
        if (vals_out) {
            // original index = perm1[perm2[i]]
            k_gather_u32<<<nblk(n), TPB, 0, st>>>(perm1.p, perm2.p, n, iota.p);   // iota reused as composed perm
            final_perm = iota.p;
        }
    } else {
        DevBuf<unsigned long long> zero; MLG_TRY(zero.alloc(n));
        CUDA_TRY(cudaMemsetAsync(zero.p, 0, (size_t)n * 8, st));
        k_zip_keys<<<nblk(n), TPB, 0, st>>>(zero.p, lo_s.p, nullptr, n, out);
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (vals_out) {
        if (vals) k_gather_u32<<<nblk(n), TPB, 0, st>>>(vals, final_perm, n, vals_out);
        else CUDA_TRY(cudaMemcpyAsync(vals_out, final_perm, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}

// level-1 table geometry from the number of distinct keys: slots per bucket and number of buckets (a power of two)
void choose_buckets(DbView& v, uint32_t nd) {
    uint32_t slots_per_bucket = 8;
    if (const char* s = getenv("MLG_BUCKET_SLOTS")) { int x = atoi(s); if (x == 4 || x == 8) slots_per_bucket = (uint32_t)x; }
    double load = slots_per_bucket * 0.3125;  // upper bound on the mean entries per bucket (2.5 of 8 slots)
    if (const char* s = getenv("MLG_BUCKET_LOAD")) { double x = atof(s); if (x > 0.01 && x <= slots_per_bucket) load = x; }
    uint32_t bbits = 1;                       // at least two buckets: the probe kernel shifts by 32 - bbits
    if (v.layout == 1) {
        // k-mers that share a minimizer share a bucket PAIR (kmer.cuh), so occupancy is clumpier than a plain hash's;
        // mean load <= 1.25 per 8-slot half keeps the overflowing halves rare
        slots_per_bucket = 8; load = 1.25; bbits = 2;
        if (const char* s = getenv("MLG_SK_LOAD")) { double x = atof(s); if (x > 0.01 && x <= 8) load = x; }
    }
    if (v.layout == 2) {
        // the bucket index only serves the exact compare of the few super-k-mers whose minimizer is in the database
        slots_per_bucket = 8; load = 2.5; bbits = 1;
        if (const char* s = getenv("MLG_MZ_LOAD")) { double x = atof(s); if (x > 0.01 && x <= 64) load = x; }
    }
    while (bbits < 31 && (double)(1ull << bbits) * load < (double)nd) ++bbits;
    v.nbuckets = 1ull << bbits; v.bbits = bbits; v.slots = slots_per_bucket;
}

}  // namespace

int mlg_db_build_device(mlg_ctx* ctx, const key128* d_keys, uint32_t G, uint32_t n, uint32_t K, const uint32_t* ks,
                        uint32_t nk, mlg_db** out) {
    if (!ctx || !out) { mlg_set_error("null argument"); return MLG_ERR_ARG; }
    if (K < 1 || K > 63) { mlg_set_error("K=%u out of range 1..63", K); return MLG_ERR_ARG; }
    if (nk < 1 || nk > MLG_MAX_KS) { mlg_set_error("nk=%u out of range 1..%d", nk, MLG_MAX_KS); return MLG_ERR_ARG; }
    for (uint32_t i = 0; i < nk; ++i)
        if (ks[i] < 1 || ks[i] > K || (i && ks[i] <= ks[i - 1])) { mlg_set_error("ks must be ascending and within 1..K"); return MLG_ERR_ARG; }
    if (K - ks[0] + 1 > 64) { mlg_set_error("K - ks[0] + 1 must be <= 64"); return MLG_ERR_ARG; }
    unsigned long long total = (unsigned long long)G * n;
    if (G == 0 || n == 0 || total >= 0x7FFFFFF0ull) { mlg_set_error("G*n=%llu out of range (1 .. 2^31-16)", total); return MLG_ERR_ARG; }
    CUDA_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->s_comp;
    cudaEvent_t e0, e1; CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventRecord(e0, st));

    mlg_db* db = new mlg_db();
    db->ctx = ctx;
    struct Guard { mlg_db* d; ~Guard() { if (d) delete d; } } guard{db};
    DbView& v = db->v;
    v.G = G; v.n = n; v.K = K; v.nk = nk;
    // Metalign's K = 60 gets the minimizer-bitmap layout (kmer.cuh); other K keep the whole-k-mer hash layout.
    // MLG_LAYOUT=0 / 1 force the whole-k-mer hash layout / the fingerprint-pair super-k-mer layout (A/B measurements).
    v.layout = (K == 60) ? 2u : 0u;
    if (const char* s = getenv("MLG_LAYOUT")) { int x = atoi(s); if (x == 0 || (x == 1 && K == 60)) v.layout = (uint32_t)x; }
    for (uint32_t i = 0; i < MLG_MAX_KS; ++i) v.ks[i] = i < nk ? ks[i] : 0;

    // 1. non-empty slots
    DevBuf<unsigned char> flag; MLG_TRY(flag.alloc(total));
    MLG_TRY(db->has_empty.alloc(G));
    CUDA_TRY(cudaMemsetAsync(db->has_empty.p, 0, G, st));
    k_mark_nonempty<<<nblk(total), TPB, 0, st>>>(d_keys, total, n, flag.p, db->has_empty.p);
    DevBuf<uint32_t> slots; MLG_TRY(slots.alloc(total));
    DevBuf<unsigned long long> d_cnt; MLG_TRY(d_cnt.alloc(1));
    {
        cub::CountingInputIterator<uint32_t> it(0);
        IsNonEmpty pred{flag.p};
        void* tmp = nullptr; size_t tb = 0;
        CUDA_TRY(cub::DeviceSelect::If(nullptr, tb, it, slots.p, d_cnt.p, (int)total, pred, st));
        CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
        cudaError_t e = cub::DeviceSelect::If(tmp, tb, it, slots.p, d_cnt.p, (int)total, pred, st);
        cudaStreamSynchronize(st); cudaFree(tmp);
        if (e != cudaSuccess) { mlg_set_error("DeviceSelect failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
    }
    unsigned long long np64 = 0;
    CUDA_TRY(cudaMemcpy(&np64, d_cnt.p, 8, cudaMemcpyDeviceToHost));
    uint32_t np = (uint32_t)np64;
    v.np = np;
    flag.release();

    // 2. P: sort by key
    MLG_TRY(db->P_key.alloc(np)); MLG_TRY(db->P_slot.alloc(np));
    {
        DevBuf<unsigned long long> hi, lo; MLG_TRY(hi.alloc(np)); MLG_TRY(lo.alloc(np));
        if (np) k_gather_split<<<nblk(np), TPB, 0, st>>>(d_keys, slots.p, np, hi.p, lo.p);
        MLG_TRY(sort_keys128(hi.p, lo.p, slots.p, np, K, db->P_key.p, db->P_slot.p, st));
    }
    slots.release();
    v.P_key = db->P_key.p; v.P_slot = db->P_slot.p;

    // 3. bucket index on the top pbits of the key
    {
        uint32_t pbits = 8;
        while (pbits < 28 && (1ull << pbits) < np) ++pbits;
        if (pbits > 2 * K) pbits = 2 * K;
        v.pbits = pbits;
        size_t m = (size_t)1 << pbits;
        MLG_TRY(db->pidx.alloc(m + 1));
        CUDA_TRY(cudaMemsetAsync(db->pidx.p, 0, (m + 1) * 4, st));
        if (np) k_pbucket_hist<<<nblk(np), TPB, 0, st>>>(db->P_key.p, np, K, pbits, db->pidx.p);
        MLG_TRY(exclusive_scan_u32(db->pidx.p, m, st));
        v.pidx = db->pidx.p;
    }

    // 4. (genome, k-prefix) classes and denominators
    MLG_TRY(db->rep.alloc((size_t)nk * total));
    CUDA_TRY(cudaMemsetAsync(db->rep.p, 0xFF, (size_t)nk * total * 4, st));
    MLG_TRY(db->den_real.alloc((size_t)G * nk));
    CUDA_TRY(cudaMemsetAsync(db->den_real.p, 0, (size_t)G * nk * 8, st));
    if (np) {
        DevBuf<uint32_t> gkey, gkey_s, iota, q2p;
        MLG_TRY(gkey.alloc(np)); MLG_TRY(gkey_s.alloc(np)); MLG_TRY(iota.alloc(np)); MLG_TRY(q2p.alloc(np));
        k_slot_to_genome<<<nblk(np), TPB, 0, st>>>(db->P_slot.p, np, n, gkey.p);
        k_iota<<<nblk(np), TPB, 0, st>>>(iota.p, np);
        int gbits = 1; while (gbits < 32 && (1ull << gbits) < G) ++gbits;
        MLG_TRY(sort_pairs_u32(gkey.p, gkey_s.p, iota.p, q2p.p, np, gbits, st));
        for (uint32_t ki = 0; ki < nk; ++ki)
            k_rep_classes<<<nblk(np), TPB, 0, st>>>(db->P_key.p, db->P_slot.p, q2p.p, np, n, K, ks[ki], ki, nk,
                                                    db->rep.p + (size_t)ki * total, db->den_real.p);
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
    }
    v.rep = db->rep.p;

    // 5. D: distinct canonical keys, hash-ordered, with the level-1 fingerprint table
    uint32_t nd = 0;
    DevBuf<unsigned long long> hsorted;
    if (np) {
        DevBuf<key128> csorted; MLG_TRY(csorted.alloc(np));
        {
            DevBuf<unsigned long long> hi, lo; MLG_TRY(hi.alloc(np)); MLG_TRY(lo.alloc(np));
            k_canon_split<<<nblk(np), TPB, 0, st>>>(db->P_key.p, np, K, hi.p, lo.p);
            MLG_TRY(sort_keys128(hi.p, lo.p, nullptr, np, K, csorted.p, nullptr, st));
        }
        DevBuf<unsigned char> head; MLG_TRY(head.alloc(np));
        k_flag_heads<<<nblk(np), TPB, 0, st>>>(csorted.p, np, head.p);
        DevBuf<key128> duniq; MLG_TRY(duniq.alloc(np));
        {
            void* tmp = nullptr; size_t tb = 0;
            CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tb, csorted.p, head.p, duniq.p, d_cnt.p, (int)np, st));
            CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
            cudaError_t e = cub::DeviceSelect::Flagged(tmp, tb, csorted.p, head.p, duniq.p, d_cnt.p, (int)np, st);
            cudaStreamSynchronize(st); cudaFree(tmp);
            if (e != cudaSuccess) { mlg_set_error("DeviceSelect failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
        }
        unsigned long long nd64 = 0;
        CUDA_TRY(cudaMemcpy(&nd64, d_cnt.p, 8, cudaMemcpyDeviceToHost));
        nd = (uint32_t)nd64;
        csorted.release();
        // how many sketch slots hold each distinct K-mer (what `kmc -ci0` of the sketches' dump counts)
        DevBuf<unsigned char> multu; MLG_TRY(multu.alloc(nd));
        {
            DevBuf<unsigned char> runlen; MLG_TRY(runlen.alloc(np));
            k_run_lengths<<<nblk(np), TPB, 0, st>>>(head.p, np, runlen.p);
            void* tmp = nullptr; size_t tb = 0;
            CUDA_TRY(cub::DeviceSelect::Flagged(nullptr, tb, runlen.p, head.p, multu.p, d_cnt.p, (int)np, st));
            CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
            cudaError_t e = cub::DeviceSelect::Flagged(tmp, tb, runlen.p, head.p, multu.p, d_cnt.p, (int)np, st);
            cudaStreamSynchronize(st); cudaFree(tmp);
            if (e != cudaSuccess) { mlg_set_error("DeviceSelect failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
        }
        head.release();
        // order by hash
        DevBuf<unsigned long long> h; DevBuf<uint32_t> iota, perm;
        MLG_TRY(h.alloc(nd)); MLG_TRY(hsorted.alloc(nd)); MLG_TRY(iota.alloc(nd)); MLG_TRY(perm.alloc(nd));
        choose_buckets(v, nd);
        k_hash_keys<<<nblk(nd), TPB, 0, st>>>(duniq.p, nd, K, v.layout, v.bbits, h.p);
        k_iota<<<nblk(nd), TPB, 0, st>>>(iota.p, nd);
        MLG_TRY(sort_pairs_u64(h.p, hsorted.p, iota.p, perm.p, nd, 0, 64, st));
        MLG_TRY(db->D_key.alloc(nd));
        k_gather_key<<<nblk(nd), TPB, 0, st>>>(duniq.p, perm.p, nd, db->D_key.p);
        MLG_TRY(db->D_mult.alloc(nd));
        k_gather_u8<<<nblk(nd), TPB, 0, st>>>(multu.p, perm.p, nd, db->D_mult.p);
        CUDA_TRY(cudaStreamSynchronize(st));
    } else {
        MLG_TRY(db->D_key.alloc(1)); MLG_TRY(hsorted.alloc(1));
    }
    v.nd = nd; v.D_key = db->D_key.p;
    {
        if (!nd) choose_buckets(v, nd);
        const uint32_t slots_per_bucket = v.slots, bbits = v.bbits;
        const unsigned long long nb = v.nbuckets;
        MLG_TRY(db->bstart.alloc(nb + 1));
        CUDA_TRY(cudaMemsetAsync(db->bstart.p, 0, (nb + 1) * 4, st));
        if (nd) k_hbucket_hist<<<nblk(nd), TPB, 0, st>>>(hsorted.p, nd, bbits, db->bstart.p);
        MLG_TRY(exclusive_scan_u32(db->bstart.p, nb, st));
        // layout 2 has no fingerprint table: its level 1 is the minimizer bitmap
        const unsigned long long t1_words = v.layout == 2 ? 8ull : nb * slots_per_bucket + 8;
        MLG_TRY(db->T1.alloc(t1_words));
        CUDA_TRY(cudaMemsetAsync(db->T1.p, 0, t1_words * 4, st));
        if (nd && v.layout != 2) k_fill_t1<<<nblk(nd), TPB, 0, st>>>(hsorted.p, nd, bbits, db->bstart.p, slots_per_bucket, v.layout, db->T1.p);
        v.bstart = db->bstart.p; v.T1 = db->T1.p;
        // One-bit-per-key prefilter sized to stay L2-resident: MLG_FILTER_MB MiB at most (default 64), 16 bits per
        // key at most; below 1.5 bits per key it would pass most probes and is left out.
        double max_mb = 64.0;
        if (const char* s = getenv("MLG_FILTER_MB")) max_mb = atof(s);
        unsigned long long nfw = 0;                       // 32-bit words
        if (nd && max_mb > 0 && v.layout == 0) {
            unsigned long long want = ((unsigned long long)nd * 16ull + 31ull) / 32ull;
            unsigned long long cap = (unsigned long long)(max_mb * 1048576.0 / 4.0);
            nfw = want < cap ? want : cap;
            if (nfw > 0xFFFFFFF0ull) nfw = 0xFFFFFFF0ull;
            if ((double)nfw * 32.0 < 1.5 * (double)nd) nfw = 0;
        }
        if (v.layout == 2) {
            // level-1 bit array: at least 32 bits per K-mer, two bits set per K-mer (density <= 1/16, false-positive rate
            // of a run ~ density^2 <= 0.4 %; every false positive costs an exact compare of ~11 windows), 2^20 .. 2^36 bits
            uint32_t fbits = 20;
            while (fbits < 36 && (1ull << fbits) < 32ull * nd) ++fbits;
            if (const char* s = getenv("MLG_MZ_FBITS")) { int x = atoi(s); if (x >= 10 && x <= 36) fbits = (uint32_t)x; }
            v.fbits = fbits;
            nfw = 1ull << (fbits - 5);
        }
        v.nfw = (uint32_t)nfw; v.F = nullptr;
        v.fk = ((double)nfw * 32.0 >= 3.0 * (double)nd) ? 2u : 1u;     // two probe bits pay off above ~3 bits per key
        if (const char* s = getenv("MLG_FILTER_K")) { int x = atoi(s); if (x == 1 || x == 2) v.fk = (uint32_t)x; }
        if (v.layout == 2) {
            MLG_TRY(db->F.alloc(nfw));
            CUDA_TRY(cudaMemsetAsync(db->F.p, 0, nfw * 4, st));
            v.F = db->F.p;
            v.n_alias = 0; v.alias_z = nullptr; v.alias_i = nullptr;
            MLG_TRY(db->alias_bloom.alloc(2048));
            CUDA_TRY(cudaMemsetAsync(db->alias_bloom.p, 0, 2048 * 4, st));
            v.alias_bloom = db->alias_bloom.p;
            if (nd) {
                CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, 8, st));
                k_fill_mzbits<<<nblk(nd), TPB, 0, st>>>(db->D_key.p, nd, K, v.fbits, db->F.p, 0, nullptr, nullptr, nullptr, d_cnt.p);
                unsigned long long na = 0;
                CUDA_TRY(cudaMemcpyAsync(&na, d_cnt.p, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                if (na) {
                    DevBuf<unsigned long long> az; DevBuf<uint32_t> ai;
                    MLG_TRY(az.alloc(na)); MLG_TRY(ai.alloc(na));
                    MLG_TRY(db->alias_z.alloc(na)); MLG_TRY(db->alias_i.alloc(na));
                    CUDA_TRY(cudaMemsetAsync(d_cnt.p, 0, 8, st));
                    k_fill_mzbits<<<nblk(nd), TPB, 0, st>>>(db->D_key.p, nd, K, v.fbits, db->F.p, 1, az.p, ai.p, db->alias_bloom.p, d_cnt.p);
                    MLG_TRY(sort_pairs_u64(az.p, db->alias_z.p, ai.p, db->alias_i.p, (uint32_t)na, 0, 64, st));
                    v.n_alias = (uint32_t)na; v.alias_z = db->alias_z.p; v.alias_i = db->alias_i.p;
                }
            }
        } else if (nfw) {
            MLG_TRY(db->F.alloc(nfw));
            CUDA_TRY(cudaMemsetAsync(db->F.p, 0, nfw * 4, st));
            k_fill_filter<<<nblk(nd), TPB, 0, st>>>(hsorted.p, nd, (uint32_t)nfw, v.fk, db->F.p);
            v.F = db->F.p;
            // pin the prefilter in L2: grow the persisting carve-out (default 24 MB of the 79 MB this part allows) and
            // put a persisting access-policy window over F on the compute stream.  MLG_L2_PERSIST=0 leaves the
            // device limits alone (the kernel's evict_last / evict_first load policies still apply).
            const char* pe = getenv("MLG_L2_PERSIST");
            if (!pe || atoi(pe) > 0) {
                cudaDeviceProp prop;
                if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
                    size_t fbytes = (size_t)nfw * 4;
                    size_t carve = fbytes < (size_t)prop.persistingL2CacheMaxSize ? fbytes : (size_t)prop.persistingL2CacheMaxSize;
                    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
                    cudaStreamAttrValue attr;
                    memset(&attr, 0, sizeof(attr));
                    attr.accessPolicyWindow.base_ptr = (void*)db->F.p;
                    attr.accessPolicyWindow.num_bytes = fbytes < (size_t)prop.accessPolicyMaxWindowSize ? fbytes : (size_t)prop.accessPolicyMaxWindowSize;
                    attr.accessPolicyWindow.hitRatio = (float)((double)carve / (double)fbytes);
                    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    if (cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
                }
            }
        }
    }
    // 6. hit records: what a present k-mer of D contributes to the per-genome table is a static function of the
    //    database, so it is expanded once here (same code as the on-the-fly kernel) and replayed per query.  Two passes:
    //    tally (sizes; the common one-slot records are complete after it) -> offsets -> fill.  Afterwards P, its bucket
    //    index and the class representatives -- 60 % of the database's footprint -- have no reader left and are released.
    //    MLG_PRECOMPUTE_HITS=0 skips all this (every query expands on the fly); MLG_HIT_CAP_WORDS=n drops the records
    //    beyond word n (their k-mers are expanded on the fly: the mixed path, for tests); MLG_KEEP_P=1 keeps P regardless.
    {
        const char* pe = getenv("MLG_PRECOMPUTE_HITS");
        if (nd && (!pe || atoi(pe) != 0)) {
            unsigned long long drop_from = ~0ull;
            if (const char* s = getenv("MLG_HIT_CAP_WORDS")) { unsigned long long x = strtoull(s, nullptr, 10); if (x >= 16) drop_from = x; }
            const unsigned long long groups = ((unsigned long long)nd + (1ull << MLG_HGROUP_SHIFT) - 1) >> MLG_HGROUP_SHIFT;
            DevBuf<uint32_t> summary;
            MLG_TRY(summary.alloc(2ull * nd)); MLG_TRY(db->hoff.alloc(nd)); MLG_TRY(db->hbase.alloc(groups + 1));
            MLG_TRY(launch_tally_hits(v, summary.p, st));
            CUDA_TRY(cudaMemsetAsync(db->hbase.p, 0, (groups + 1) * 8, st));
            MLG_TRY(launch_hit_group_sums(summary.p, nd, db->hbase.p, st));
            {
                void* tmp = nullptr; size_t tb = 0;
                CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, db->hbase.p, db->hbase.p, (int)(groups + 1), st));
                CUDA_TRY(cudaMalloc(&tmp, tb ? tb : 1));
                cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, db->hbase.p, db->hbase.p, (int)(groups + 1), st);
                cudaStreamSynchronize(st); cudaFree(tmp);
                if (e != cudaSuccess) { mlg_set_error("DeviceScan failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
            }
            MLG_TRY(launch_hit_offsets(summary.p, nd, db->hoff.p, st));
            unsigned long long total_words = 0;
            CUDA_TRY(cudaMemcpyAsync(&total_words, db->hbase.p + groups, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaGetLastError());
            const unsigned long long keep_words = total_words < drop_from ? total_words : drop_from;
            MLG_TRY(db->hits.alloc(keep_words + 2 + MLG_MAX_KS));
            MLG_TRY(launch_fill_hits(v, summary.p, db->hoff.p, db->hbase.p, db->hits.p, drop_from, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            CUDA_TRY(cudaGetLastError());
            db->hit_words = keep_words;
            const char* kp = getenv("MLG_KEEP_P");
            if (drop_from == ~0ull && !(kp && atoi(kp) != 0)) {
                db->P_key.release(); db->P_slot.release(); db->pidx.release(); db->rep.release();
                v.P_key = nullptr; v.P_slot = nullptr; v.pidx = nullptr; v.rep = nullptr;
                db->p_dropped = true;
            }
        }
    }
    CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); db->build_ms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    guard.d = nullptr;
    mlg_pool_trim(ctx->device);   // the build's multi-GB temporaries should not stay cached
    *out = db;
    return MLG_OK;
}
