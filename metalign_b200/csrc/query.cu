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
//   k_popcount_table per-(genome, k) popcount of the hit bitmap, accumulated in shared memory with
//                   warp-aggregated atomics (match_any + redux), then flushed to global.
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

__device__ __forceinline__ void mark_hit(const DbView& db, uint32_t* hitbits, unsigned long long words_per_k, uint32_t ki, uint32_t e) {
    const unsigned long long total = (unsigned long long)db.G * db.n;
    const uint32_t slot = db.P_slot[e];
    const uint32_t r = db.rep[(unsigned long long)ki * total + slot];
    atomicOr(&hitbits[(unsigned long long)ki * words_per_k + (r >> 5)], 1u << (r & 31u));
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_expand_hits(DbView db, const uint32_t* __restrict__ present,
                                                                    uint32_t n_present, int gate_none, uint32_t* hitbits,
                                                                    unsigned long long words_per_k) {
    __shared__ uint32_t s_flo[WARPS_PER_CTA][MAX_OFF], s_fcnt[WARPS_PER_CTA][MAX_OFF];
    __shared__ uint32_t s_rlo[WARPS_PER_CTA][MAX_OFF], s_rcnt[WARPS_PER_CTA][MAX_OFF];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const unsigned K = db.K, k0 = db.ks[0], noff = K - k0 + 1;
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
            if (fc) { for (uint32_t e = fl; e < fl + fc; ++e) mark_hit(db, hitbits, words_per_k, 0, e); }
            else    { for (uint32_t e = rl; e < rl + rc; ++e) mark_hit(db, hitbits, words_per_k, 0, e); }
            const bool possible = gate_none || fc || rc;
            if (!possible) continue;
            for (uint32_t ki = 1; ki < db.nk; ++ki) {
                const unsigned k = db.ks[ki];
                if (o + k > K) continue;
                const key128 wk = key_sub(x, K, o, k);
                bool any = false;
                for (uint32_t e = fl; e < fl + fc; ++e)
                    if (key_eq(key_prefix(db.P_key[e], K, k), wk)) { mark_hit(db, hitbits, words_per_k, ki, e); any = true; }
                if (!any) {
                    // rc(wk) starts with the reverse complement of the LAST k0 bases of wk: offset o + k - k0
                    const unsigned o2 = o + k - k0;
                    const uint32_t rl2 = s_rlo[warp][o2], rc2 = s_rcnt[warp][o2];
                    const key128 rk = key_rc(wk, k);
                    for (uint32_t e = rl2; e < rl2 + rc2; ++e)
                        if (key_eq(key_prefix(db.P_key[e], K, k), rk)) mark_hit(db, hitbits, words_per_k, ki, e);
                }
            }
        }
        __syncwarp();
    }
}

constexpr int POP_TPB = 256;
constexpr int POP_CAP = 1024;   // genomes one CTA can accumulate in shared memory

__global__ void __launch_bounds__(POP_TPB) k_popcount_table(const uint32_t* __restrict__ hitbits, unsigned long long words_per_k,
                                                            uint32_t G, uint32_t n, uint32_t nk, unsigned long long* num) {
    __shared__ unsigned int s_cnt[POP_CAP];
    const uint32_t ki = blockIdx.y;
    const unsigned long long w0 = (unsigned long long)blockIdx.x * POP_TPB;
    const unsigned long long total = (unsigned long long)G * n;
    const unsigned long long slot_a = w0 * 32ull;
    unsigned long long slot_b = slot_a + (unsigned long long)POP_TPB * 32ull;   // exclusive
    if (slot_b > total) slot_b = total;
    if (slot_a >= total) return;
    const uint32_t gA = (uint32_t)(slot_a / n), gB = (uint32_t)((slot_b - 1) / n);
    const bool use_smem = (gB - gA + 1) <= POP_CAP;
    if (use_smem) for (unsigned i = threadIdx.x; i <= gB - gA; i += POP_TPB) s_cnt[i] = 0;
    __syncthreads();

    const unsigned long long wi = w0 + threadIdx.x;
    uint32_t word = 0;
    unsigned long long s0 = wi * 32ull;
    if (wi < words_per_k && s0 < total) word = hitbits[(unsigned long long)ki * words_per_k + wi];
    if (s0 >= total) s0 = total - 1;   // inactive lanes still take part in the warp collectives
    const uint32_t g_first = (uint32_t)(s0 / n);
    // bits of this word that belong to g_first
    unsigned long long g_end = ((unsigned long long)g_first + 1) * n;     // first slot of the next genome
    uint32_t nbits = (g_end - s0) >= 32ull ? 32u : (uint32_t)(g_end - s0);
    uint32_t m_first = nbits >= 32u ? 0xFFFFFFFFu : ((1u << nbits) - 1u);
    uint32_t c_first = __popc(word & m_first);
    // warp-aggregated add for the first (usually only) genome of the word
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, g_first);
    const unsigned sum = __reduce_add_sync(peers, c_first);
    const bool leader = (threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1);
    if (leader && sum) {
        if (use_smem) atomicAdd(&s_cnt[g_first - gA], sum);
        else atomicAdd(&num[(unsigned long long)g_first * nk + ki], (unsigned long long)sum);
    }
    // a word can straddle further genomes (always when n < 32)
    uint32_t rest = nbits >= 32u ? 0u : (word >> nbits);
    uint32_t g = g_first + 1;
    uint32_t left = 32u - nbits;
    while (left > 0 && g < G) {
        uint32_t take = n < left ? n : left;
        uint32_t m = take >= 32u ? 0xFFFFFFFFu : ((1u << take) - 1u);
        uint32_t c = __popc(rest & m);
        if (c) {
            if (use_smem) atomicAdd(&s_cnt[g - gA], c);
            else atomicAdd(&num[(unsigned long long)g * nk + ki], (unsigned long long)c);
        }
        rest = take >= 32u ? 0u : (rest >> take);
        left -= take; ++g;
    }
    __syncthreads();
    if (use_smem)
        for (unsigned i = threadIdx.x; i <= gB - gA; i += POP_TPB)
            if (s_cnt[i]) atomicAdd(&num[(unsigned long long)(gA + i) * nk + ki], (unsigned long long)s_cnt[i]);
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
__global__ void k_count_present(const unsigned char* cnt8, uint32_t nd, uint32_t ci_min, unsigned long long* out) {
    unsigned long long c = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < nd;
         i += (unsigned long long)gridDim.x * blockDim.x)
        c += cnt8[i] >= ci_min;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31u) == 0 && c) atomicAdd(out, c);
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
int launch_count_present(const unsigned char* cnt8, uint32_t nd, uint32_t ci_min, unsigned long long* d_count, cudaStream_t st) {
    CUDA_TRY(cudaMemsetAsync(d_count, 0, 8, st));
    if (!nd) return MLG_OK;
    unsigned grid = (unsigned)(((unsigned long long)nd + 256ull * 16 - 1) / (256ull * 16));
    if (grid > 148u * 8u) grid = 148u * 8u;
    k_count_present<<<grid, 256, 0, st>>>(cnt8, nd, ci_min, d_count);
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
int launch_expand_hits(const DbView& db, const uint32_t* present, uint32_t n_present, int gate_none, uint32_t* hitbits,
                       unsigned long long words_per_k, cudaStream_t st) {
    if (!n_present) return MLG_OK;
    unsigned long long want = ((unsigned long long)n_present + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    unsigned grid = (unsigned)(want < 148ull * 16ull ? want : 148ull * 16ull);
    k_expand_hits<<<grid, WARPS_PER_CTA * 32, 0, st>>>(db, present, n_present, gate_none, hitbits, words_per_k);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_popcount_table(const uint32_t* hitbits, unsigned long long words_per_k, uint32_t G, uint32_t n, uint32_t nk,
                          unsigned long long* num, cudaStream_t st) {
    dim3 grid((unsigned)((words_per_k + POP_TPB - 1) / POP_TPB), nk);
    k_popcount_table<<<grid, POP_TPB, 0, st>>>(hitbits, words_per_k, G, n, nk, num);
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
