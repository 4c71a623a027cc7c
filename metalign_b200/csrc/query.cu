// K2 / K3: from the set I of present database k-mers to the per-genome, per-k containment table.
// Replaces the per-record loop of CMash's StreamingQueryDNADatabase.py as Metalign calls it
// (scripts/select_db.py:73-76: `... 30-60-10 -c 0 -r 1000000 -v -f <bf> --sensitive`) up to the
// DataFrame tail: SURVEY.md 3.3 steps R4 (prefix matching, forward first / reverse complement only if
// the forward lookup is empty, smallest-k prefilter gate) and R5 (distinct hit prefixes per genome, k).
//
//   k_expand_hits   one warp per present k-mer x, one lane per offset: the ks[0]-mer at every offset is
//                   looked up forward and reverse-complemented in the sorted array P (bucketed binary
//                   search on the top key bits; a k-prefix is a key range, so every k is served by the
//                   same array); longer k's refine those ranges.  A hit ORs one bit: (k, class
//                   representative slot) -- set semantics, never a sum.
//                   The lane that sets a bit FIRST (atomicOr returned it clear) also adds 1 to num[genome][k]:
//                   the per-genome table is complete when this kernel ends.  (A separate popcount pass
//                   over the whole G*n-bit bitmap cost 0.17 ms at 2e5 genomes for ~2e5 set bits.)
//   k_finalize      containment = num / den in IEEE double where num > 0.
#include "mlg_internal.h"

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr int MAX_OFF = 64;

// entries of P whose leading k bases equal v: first index and count
__device__ __forceinline__ void prefix_range(const DbView& db, const key128& v, unsigned k, uint32_t& lo_out, uint32_t& cnt_out) {
    const unsigned K = db.K;
    const key128 lo_key = key_shl(v, 2 * (K - k));
    const uint32_t b = (uint32_t)key_shr(lo_key, 2 * K - db.pbits).lo;
    uint32_t lo = db.pidx[b], hi = db.pidx[b + 1];
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (key_lt(db.P_key[mid], lo_key)) lo = mid + 1; else hi = mid;
    }
    uint32_t e = lo;
    while (e < db.np && key_eq(key_prefix(db.P_key[e], K, k), v)) ++e;
    lo_out = lo; cnt_out = e - lo;
}

struct HitSink {
    uint32_t* hitbits;
    unsigned long long words_per_k;
    unsigned long long* num;      // G * nk
};
__device__ __forceinline__ void mark_hit(const DbView& db, const HitSink& hs, uint32_t ki, uint32_t e) {
    const unsigned long long total = (unsigned long long)db.G * db.n;
    const uint32_t slot = db.P_slot[e];
    const uint32_t r = db.rep[(unsigned long long)ki * total + slot];
    const uint32_t bit = 1u << (r & 31u);
    const uint32_t old = atomicOr(&hs.hitbits[(unsigned long long)ki * hs.words_per_k + (r >> 5)], bit);
    if (!(old & bit)) atomicAdd(&hs.num[(unsigned long long)(r / db.n) * db.nk + ki], 1ull);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_expand_hits(DbView db, const uint32_t* __restrict__ present,
                                                                    const unsigned long long* __restrict__ d_n_present,
                                                                    int gate_none, HitSink hs) {
    __shared__ uint32_t s_flo[WARPS_PER_CTA][MAX_OFF], s_fcnt[WARPS_PER_CTA][MAX_OFF];
    __shared__ uint32_t s_rlo[WARPS_PER_CTA][MAX_OFF], s_rcnt[WARPS_PER_CTA][MAX_OFF];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned K = db.K, k0 = db.ks[0], noff = K - k0 + 1;
    const unsigned long long n_present = *d_n_present;
    for (unsigned long long xi = (unsigned long long)blockIdx.x * WARPS_PER_CTA + warp; xi < n_present;
         xi += (unsigned long long)gridDim.x * WARPS_PER_CTA) {
        const key128 x = db.D_key[present[xi]];
        for (unsigned o = lane; o < noff; o += 32) {
            const key128 w = key_sub(x, K, o, k0);
            uint32_t lo, cnt;
            prefix_range(db, w, k0, lo, cnt);
            s_flo[warp][o] = lo; s_fcnt[warp][o] = cnt;
            prefix_range(db, key_rc(w, k0), k0, lo, cnt);
            s_rlo[warp][o] = lo; s_rcnt[warp][o] = cnt;
        }
        __syncwarp();
        for (unsigned o = lane; o < noff; o += 32) {
            const uint32_t fl = s_flo[warp][o], fc = s_fcnt[warp][o];
            const uint32_t rl = s_rlo[warp][o], rc = s_rcnt[warp][o];
            // smallest k: forward first, reverse complement only if forward is empty
            if (fc) { for (uint32_t e = fl; e < fl + fc; ++e) mark_hit(db, hs, 0, e); }
            else    { for (uint32_t e = rl; e < rl + rc; ++e) mark_hit(db, hs, 0, e); }
            const bool possible = gate_none || fc || rc;
            if (!possible) continue;
            for (uint32_t ki = 1; ki < db.nk; ++ki) {
                const unsigned k = db.ks[ki];
                if (o + k > K) continue;
                const key128 wk = key_sub(x, K, o, k);
                bool any = false;
                for (uint32_t e = fl; e < fl + fc; ++e)
                    if (key_eq(key_prefix(db.P_key[e], K, k), wk)) { mark_hit(db, hs, ki, e); any = true; }
                if (!any) {
                    // rc(wk) starts with the reverse complement of the LAST k0 bases of wk: offset o + k - k0
                    const unsigned o2 = o + k - k0;
                    const uint32_t rl2 = s_rlo[warp][o2], rc2 = s_rcnt[warp][o2];
                    const key128 rk = key_rc(wk, k);
                    for (uint32_t e = rl2; e < rl2 + rc2; ++e)
                        if (key_eq(key_prefix(db.P_key[e], K, k), rk)) mark_hit(db, hs, ki, e);
                }
            }
        }
        __syncwarp();
    }
}

__global__ void k_finalize(const unsigned long long* num, const long long* den_real, const unsigned char* has_empty, uint32_t G,
                           uint32_t nk, int count_empty, long long* out_num, long long* out_den, double* out_ci) {
    unsigned long long c = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (c >= (unsigned long long)G * nk) return;
    uint32_t g = (uint32_t)(c / nk);
    long long nu = (long long)num[c];
    long long de = den_real[c] + ((count_empty && has_empty[g]) ? 1 : 0);
    out_num[c] = nu; out_den[c] = de;
    out_ci[c] = nu > 0 ? (double)nu / (double)de : 0.0;
}

__global__ void k_clamp_counts(uint32_t* cnt_words, unsigned long long nwords, uint32_t ci_min) {
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    uint32_t w = cnt_words[i], r = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        uint32_t c = (w >> (8 * b)) & 0xFFu;
        r |= (c > ci_min ? ci_min : c) << (8 * b);
    }
    cnt_words[i] = r;
}
__global__ void k_compact_present(const unsigned char* cnt8, uint32_t nd, uint32_t ci_min, uint32_t* out, unsigned long long* cursor) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long nd_round = ((unsigned long long)nd + 31ull) & ~31ull;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nd_round; i += stride) {
        const bool p = i < nd && cnt8[i] >= ci_min;
        const unsigned ballot = __ballot_sync(0xFFFFFFFFu, p);
        if (ballot == 0) continue;
        const unsigned lane = threadIdx.x & 31u;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(ballot));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (p) out[base + __popc(ballot & ((1u << lane) - 1u))] = (uint32_t)i;
    }
}
// N runs -> N mask.  One thread per (start, length) run; bit i of the mask lives in byte i/8, bit 7-(i%8).
__global__ void k_scatter_nruns(const uint32_t* __restrict__ runs, unsigned long long n_runs, unsigned long long nbases,
                                uint32_t* mask_words) {
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (t >= n_runs) return;
    unsigned long long a = runs[2 * t], b = a + runs[2 * t + 1];
    if (b > nbases) b = nbases;
    while (a < b) {
        const unsigned long long byte = a >> 3;
        const unsigned first = (unsigned)(a & 7ull);
        unsigned long long stop = (byte + 1) << 3;
        if (stop > b) stop = b;
        const unsigned cnt = (unsigned)(stop - a);                       // bits first .. first+cnt-1, MSB first
        const uint32_t m = ((0xFFu >> first) & (0xFFu << (8u - first - cnt))) & 0xFFu;
        atomicOr(&mask_words[byte >> 2], m << (8u * (unsigned)(byte & 3ull)));
        a = stop;
    }
}
__global__ void k_gather_present_keys(const key128* D_key, const uint32_t* present, uint32_t n_present, key128* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_present) out[i] = D_key[present[i]];
}

}  // namespace

int launch_clamp_counts(unsigned char* cnt8, uint32_t nd, uint32_t ci_min, cudaStream_t st) {
    unsigned long long nwords = ((unsigned long long)nd + 3) / 4;
    if (!nwords) return MLG_OK;
    k_clamp_counts<<<(unsigned)((nwords + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint32_t*>(cnt8), nwords, ci_min);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_compact_present(const unsigned char* cnt8, uint32_t nd, uint32_t ci_min, uint32_t* out, unsigned long long* d_cursor,
                           cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(d_cursor, 0, 8, st));
    if (!nd) return MLG_OK;
    unsigned grid = (unsigned)(((unsigned long long)nd + 256ull * 16 - 1) / (256ull * 16));
    if (grid > 148u * 8u) grid = 148u * 8u;
    k_compact_present<<<grid, 256, 0, st>>>(cnt8, nd, ci_min, out, d_cursor);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_expand_hits(const DbView& db, const uint32_t* present, const unsigned long long* d_n_present, int gate_none,
                       uint32_t* hitbits, unsigned long long words_per_k, unsigned long long* num, cudaStream_t st) {
    // persistent grid: |I| is only known on the device (no host round trip between the probe and this kernel)
    HitSink hs{hitbits, words_per_k, num};
    k_expand_hits<<<148u * 16u, WARPS_PER_CTA * 32, 0, st>>>(db, present, d_n_present, gate_none, hs);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_scatter_nruns(const uint32_t* d_runs, unsigned long long n_runs, unsigned long long nbases, unsigned char* nmask,
                         cudaStream_t st) {
    if (!n_runs) return MLG_OK;
    k_scatter_nruns<<<(unsigned)((n_runs + 255) / 256), 256, 0, st>>>(d_runs, n_runs, nbases, reinterpret_cast<uint32_t*>(nmask));
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_finalize(const unsigned long long* num, const long long* den_real, const unsigned char* has_empty, uint32_t G,
                    uint32_t nk, int count_empty, long long* out_num, long long* out_den, double* out_ci, cudaStream_t st) {
    unsigned long long cells = (unsigned long long)G * nk;
    k_finalize<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(num, den_real, has_empty, G, nk, count_empty, out_num, out_den, out_ci);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_gather_keys(const key128* D_key, const uint32_t* present, uint32_t n_present, key128* out, cudaStream_t st) {
    if (!n_present) return MLG_OK;
    k_gather_present_keys<<<(n_present + 255) / 256, 256, 0, st>>>(D_key, present, n_present, out);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
