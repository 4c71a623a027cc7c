// K1: decode + canonicalise + probe + count.   Replaces `kmc -k60 -ci2 -cs3` followed by
// `kmc_tools simple <db> <reads> intersect` (scripts/select_db.py:50-56): every N-free K-long window of every
// read is counted under its canonical form (min of forward / reverse complement, A<C<G<T) -- but only for
// k-mers of the database set D, which is all the intersection keeps.
//
// Mapping: one lane per read, WMAX = 96 window starts at a time ("segment"; a 150-base read is one segment of
// 91 windows), so no lane ever spends work on a window that straddles two reads.
//   per segment (once):  the lane gathers its 160 bases and 160 N bits from the packed stream into registers,
//                        top-aligned, and builds the reverse complement of the whole segment (brev + bit-pair
//                        swap); window validity (no N inside) is a bit mask made by shift-or smearing;
//   per window:          forward and reverse-complement k-mers are funnel-shift extractions from those
//                        register arrays (no loop-carried dependency); their word-wise sum is a strand-symmetric
//                        digest, hashed once (kmer.cuh);
//                        level 0: ONE bit of an L2-resident bit array (prefilter); level 1, only if that bit is
//                        set: ONE 32-byte bucket of 31-bit fingerprints from HBM;
//   rare:                fingerprint match or overflowed bucket -> the canonical key is queued per warp and the
//                        queue is drained 32 at a time, one candidate per lane (exact compare against the bucket's
//                        run of D, saturating 8-bit counter bump by 32-bit CAS).
// The tile of 256 reads a CTA works on is staged global -> shared by TMA bulk copies (cp.async.bulk + mbarrier,
// double buffered) when it fits; lanes then gather from shared memory.  Tiles that do not fit (very long reads)
// are gathered straight from global memory.
#include "probe_common.cuh"

namespace {

#ifndef K1_GROUP
#define K1_GROUP 4                      // windows whose loads are in flight together, per lane
#endif
#ifndef K1_MINCTAS
#define K1_MINCTAS 2                    // __launch_bounds__ minimum CTAs per SM (caps registers at 65536 / (256 * K1_MINCTAS))
#endif
constexpr unsigned QCAP = 64;           // per-warp queue of exact-path candidates (drained at >= 32)
constexpr unsigned STAGE_B = 16384 + 64;   // staged bytes of packed bases per tile (256 reads x 250 bases fit)
constexpr unsigned STAGE_M = 8192 + 64;    // staged bytes of N mask per tile

template <int SLOTS> struct BucketVec;
template <> struct BucketVec<8> { uint32_t w[8]; };
template <> struct BucketVec<4> { uint32_t w[4]; };

__device__ __forceinline__ void load_bucket(const uint32_t* T1, unsigned long long b, BucketVec<8>& v, unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
                 : "l"(T1 + b * 8), "l"(pol));
}
__device__ __forceinline__ void load_bucket(const uint32_t* T1, unsigned long long b, BucketVec<4>& v, unsigned long long pol) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.b32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(T1 + b * 4), "l"(pol));
}
__device__ __forceinline__ uint32_t load_filter(const uint32_t* p, unsigned long long pol) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
// does this bucket need the exact path for fingerprint fp?  (bit 31 of word 0 = bucket overflowed)
template <int SLOTS>
__device__ __forceinline__ bool bucket_candidate(const BucketVec<SLOTS>& v, uint32_t fp) {
    bool hit = ((v.w[0] & 0x7FFFFFFFu) == fp) | ((int)v.w[0] < 0);
#pragma unroll
    for (int s = 1; s < SLOTS; ++s) hit |= (v.w[s] == fp);
    return hit;
}

// exact path: compare the full key against the bucket's run of D, bump the counter on a match
__device__ __forceinline__ void probe_exact(const DbView& db, const CountSink& cs, unsigned long long khi, unsigned long long klo) {
    key128 c; c.hi = khi; c.lo = klo;
    const unsigned long long bucket = hash_bucket(key_hash(c, db.K), db.bbits);
    uint32_t s = db.bstart[bucket], e = db.bstart[bucket + 1];
    for (uint32_t i = s; i < e; ++i) {
        key128 d = db.D_key[i];
        if (d.hi == khi && d.lo == klo) { bump_counter(cs, i); return; }
    }
}

// The exact path is rare (true hits, fingerprint collisions, overflowed buckets) but made of dependent DRAM
// loads; taken inline it would serialise a whole warp on one lane.  Candidates are queued per warp in shared
// memory and drained 32 at a time, one candidate per lane.
struct WarpQueue {
    unsigned long long hi[WARPS][QCAP];
    unsigned long long lo[WARPS][QCAP];
    unsigned n[WARPS];
};
__device__ __noinline__ void queue_drain(WarpQueue& q, unsigned warp, unsigned lane, const DbView& db, const CountSink& cnt8) {
    __syncwarp();
    const unsigned n = q.n[warp];
    for (unsigned i = lane; i < n; i += 32) probe_exact(db, cnt8, q.hi[warp][i], q.lo[warp][i]);
    __syncwarp();
    if (lane == 0) q.n[warp] = 0;
    __syncwarp();
}
// all 32 lanes call this with the ballot of candidate lanes (non-zero); candidate lanes append their key
__device__ __forceinline__ void queue_push(WarpQueue& q, unsigned warp, unsigned lane, unsigned ballot, bool cand,
                                           unsigned long long khi, unsigned long long klo, const DbView& db, const CountSink& cnt8) {
    const unsigned base = q.n[warp];
    if (cand) {
        const unsigned i = base + __popc(ballot & ((1u << lane) - 1u));
        q.hi[warp][i] = khi; q.lo[warp][i] = klo;
    }
    __syncwarp();
    const unsigned total = base + __popc(ballot);
    if (lane == 0) q.n[warp] = total;
    __syncwarp();
    if (total >= 32) queue_drain(q, warp, lane, db, cnt8);
}

struct SharedStage {
    __align__(16) unsigned char b[2][STAGE_B];
    __align__(16) unsigned char m[2][STAGE_M];
};

template <int SLOTS, bool HAS_NMASK, int KT, bool USE_FILTER>
__global__ void __launch_bounds__(RT, K1_MINCTAS) k1_decode_canon_probe(ProbeArgs a, DbView db) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedStage& stg = *reinterpret_cast<SharedStage*>(smem_raw);
    WarpQueue& wq = *reinterpret_cast<WarpQueue*>(smem_raw + sizeof(SharedStage));
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ unsigned long long s_total;
    // per stage: first staged 8-byte word of bases / of N mask, and whether the tile was staged at all
    __shared__ unsigned long long s_bw0[2], s_mw0[2];
    __shared__ unsigned s_staged[2];

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned K = KT ? (unsigned)KT : db.K;
    const unsigned long long nreads = a.r_end - a.r_begin;
    const unsigned long long ntiles = (nreads + RT - 1) / RT;
    const unsigned bshift = 32u - db.bbits;      // 1 <= bbits <= 31
    const unsigned long long pol_stream = policy_evict_first(), pol_keep = policy_evict_last();
    const CountSink sink{a.cnt8, a.present, a.n_present, a.touched, a.ci_min};

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_total = 0;
    }
    if (lane == 0) wq.n[warp] = 0;
    __syncthreads();

    // ---- producer side (thread 0): stage the stream range of one tile, if it fits
    auto issue = [&](unsigned stage, unsigned long long t) {
        const unsigned long long r0 = a.r_begin + t * RT;
        const unsigned long long r1 = (r0 + RT < a.r_end) ? r0 + RT : a.r_end;
        const unsigned long long p0 = a.off ? a.off[r0] : r0 * (unsigned long long)a.read_len;
        const unsigned long long p1 = a.off ? a.off[r1] : r1 * (unsigned long long)a.read_len;
        // bases: 8-byte words [p0/32, (p1+31)/32 + 6) rounded to 16 bytes; lanes read up to 6 words past their start
        unsigned long long bw0 = (p0 >> 5) & ~1ull;
        unsigned long long bw1 = ((p1 + 31) >> 5) + 6;
        if (bw1 > a.base_words) bw1 = a.base_words;
        bw1 = (bw1 + 1) & ~1ull;                                   // base_words is even (buffers are multiples of 16 bytes)
        unsigned long long mw0 = (p0 >> 6) & ~1ull;
        unsigned long long mw1 = ((p1 + 63) >> 6) + 4;
        if (HAS_NMASK) { if (mw1 > a.nmask_words) mw1 = a.nmask_words; mw1 = (mw1 + 1) & ~1ull; }
        const unsigned long long bytes_b = (bw1 - bw0) * 8ull, bytes_m = HAS_NMASK ? (mw1 - mw0) * 8ull : 0ull;
        const bool fits = bw1 > bw0 && bytes_b <= STAGE_B && bytes_m <= STAGE_M;
        s_bw0[stage] = bw0; s_mw0[stage] = mw0; s_staged[stage] = fits ? 1u : 0u;
        if (fits) {
            mbar_expect_tx(&mbar[stage], (uint32_t)(bytes_b + bytes_m));
            bulk_g2s(&stg.b[stage][0], a.bases + bw0, (uint32_t)bytes_b, &mbar[stage], pol_stream);
            if (HAS_NMASK && bytes_m) bulk_g2s(&stg.m[stage][0], a.nmask + mw0, (uint32_t)bytes_m, &mbar[stage], pol_stream);
        } else {
            mbar_arrive(&mbar[stage]);                            // nothing to wait for: lanes gather from global
        }
    };

    // top-aligned masks of a 2K-bit value in four 32-bit words (word 3 most significant)
    uint32_t km[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int hi_bit = 32 * (j + 1), lo_valid = 128 - 2 * (int)K;   // bits >= lo_valid are k-mer bits
        km[j] = lo_valid <= 32 * j ? 0xFFFFFFFFu : (lo_valid >= hi_bit ? 0u : (0xFFFFFFFFu << (lo_valid - 32 * j)));
    }
    unsigned long long my_valid = 0;

    unsigned it = 0;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const unsigned stage = it & 1u, parity = (it >> 1) & 1u;
        if (it == 0 && tid == 0) {
            issue(0, t);
            if (t + gridDim.x < ntiles) issue(1, t + gridDim.x);
        }
        __syncthreads();                       // s_bw0 / s_staged of this stage are visible
        mbar_wait(&mbar[stage], parity);

        const unsigned long long r = a.r_begin + t * RT + tid;
        const bool active = r < a.r_end;
        unsigned long long R0 = 0, R1 = 0;
        if (active) {
            R0 = a.off ? a.off[r] : r * (unsigned long long)a.read_len;
            R1 = a.off ? a.off[r + 1] : R0 + a.read_len;
        }
        const unsigned long long len = R1 - R0;
        const unsigned long long nw = len >= K ? len - K + 1 : 0ull;
        const unsigned nseg = (unsigned)((nw + WMAX - 1) / WMAX);
        const unsigned max_seg = __reduce_max_sync(0xFFFFFFFFu, nseg);
        const bool staged = s_staged[stage] != 0;
        const unsigned long long* bsrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.b[stage][0]) - s_bw0[stage] : a.bases;
        const unsigned long long* msrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.m[stage][0]) - s_mw0[stage] : a.nmask;

        for (unsigned seg = 0; seg < max_seg; ++seg) {
            const unsigned c = seg < nseg ? (unsigned)((nw - (unsigned long long)seg * WMAX) < WMAX ? (nw - (unsigned long long)seg * WMAX) : WMAX) : 0u;
            const unsigned long long s = R0 + (unsigned long long)seg * WMAX;

            // ---- gather 160 bases (and N bits) starting at stream base s, top-aligned, into registers
            uint32_t loc[SEGW];
            uint32_t nl[5];
            segment_load<HAS_NMASK>(loc, nl, c, s, bsrc, msrc, a.base_words, a.nmask_words);
            uint32_t v0, v1, v2;
            segment_valid<HAS_NMASK>(nl, c, K, v0, v1, v2);
            my_valid += __popc(v0) + __popc(v1) + __popc(v2);
            if (__all_sync(0xFFFFFFFFu, (v0 | v1 | v2) == 0u)) continue;

            // ---- reverse complement of the whole segment, aligned so that the reverse complement of window i
            //      starts at base (WMAX-1-i) of it, whatever K is
            uint32_t rcl[SEGW];
            {
                uint32_t t160[SEGW + 1];
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) t160[k] = rev2_32(~loc[SEGW - 1 - k]);
                t160[SEGW] = 0;
                const unsigned shl = 2u * (160u - (WMAX + K - 1u));      // bits to drop at the front
                for (unsigned wsh = 0; wsh < (shl >> 5); ++wsh) {
#pragma unroll
                    for (int k = 0; k < (int)SEGW; ++k) t160[k] = t160[k + 1];
                }
                const unsigned bs = shl & 31u;
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) rcl[k] = fsl(t160[k], t160[k + 1], bs);
            }

            // ---- 6 blocks of 16 windows; the register windows loc[0..4] / rcl[5..9] slide by one word per block
#pragma unroll 1
            for (int blk = 0; blk < (int)(WMAX / 16); ++blk) {
                const uint32_t vb = v0 & 0xFFFF0000u;            // validity of the 16 windows of this block, MSB first
                v0 = fsl(v0, v1, 16); v1 = fsl(v1, v2, 16); v2 <<= 16;
                if (!__all_sync(0xFFFFFFFFu, vb == 0u)) {
                    constexpr int GROUP = K1_GROUP;
#pragma unroll
                    for (int g0 = 0; g0 < 16; g0 += GROUP) {
                        unsigned long long h[GROUP];
                        bool ok[GROUP];
#pragma unroll
                        for (int j = 0; j < GROUP; ++j) {
                            const int tt = g0 + j;
                            // forward k-mer of window tt: bases [tt, tt+K) of loc[0..4]
                            const uint32_t f3 = fsl(loc[0], loc[1], 2 * tt) & km[3];
                            const uint32_t f2 = fsl(loc[1], loc[2], 2 * tt) & km[2];
                            const uint32_t f1 = fsl(loc[2], loc[3], 2 * tt) & km[1];
                            const uint32_t f0 = fsl(loc[3], loc[4], 2 * tt) & km[0];
                            // reverse complement: bases [15-tt, 15-tt+K) of rcl[5..9]
                            const uint32_t g3 = fsl(rcl[5], rcl[6], 2 * (15 - tt)) & km[3];
                            const uint32_t g2 = fsl(rcl[6], rcl[7], 2 * (15 - tt)) & km[2];
                            const uint32_t g1 = fsl(rcl[7], rcl[8], 2 * (15 - tt)) & km[1];
                            const uint32_t gz = fsl(rcl[8], rcl[9], 2 * (15 - tt)) & km[0];
                            ok[j] = (vb >> (31 - tt)) & 1u;
                            h[j] = hash_digest(f3 + g3, f2 + g2, f1 + g1, f0 + gz);
                        }
                        if (USE_FILTER) {
                            uint32_t fw[GROUP];
#pragma unroll
                            for (int j = 0; j < GROUP; ++j) {
                                fw[j] = 0u;
                                if (ok[j]) fw[j] = load_filter(db.F + filter_word(h[j], db.nfw), pol_keep);
                            }
#pragma unroll
                            for (int j = 0; j < GROUP; ++j) {
                                const uint32_t m = filter_mask(h[j], db.fk);
                                ok[j] = ok[j] & ((fw[j] & m) == m);
                            }
                        }
                        BucketVec<SLOTS> vec[GROUP];
#pragma unroll
                        for (int j = 0; j < GROUP; ++j)
                            if (ok[j]) load_bucket(db.T1, (uint32_t)(h[j] >> 32) >> bshift, vec[j], pol_stream);
                        bool cand[GROUP];
                        bool any = false;
#pragma unroll
                        for (int j = 0; j < GROUP; ++j) {
                            cand[j] = ok[j] & bucket_candidate<SLOTS>(vec[j], hash_fp(h[j]));   // no short circuit: stay branch-free
                            any |= cand[j];
                        }
                        if (__any_sync(0xFFFFFFFFu, any)) {
#pragma unroll
                            for (int j = 0; j < GROUP; ++j) {
                                const unsigned ballot = __ballot_sync(0xFFFFFFFFu, cand[j]);
                                if (ballot) {
                                    // rebuild the two k-mers (not kept alive across the loads);
                                    // canonical = the smaller of the two (top-aligned order == numeric order)
                                    const int tt = g0 + j;
                                    key128 F, G;
                                    F.hi = ((unsigned long long)(fsl(loc[0], loc[1], 2 * tt) & km[3]) << 32) | (fsl(loc[1], loc[2], 2 * tt) & km[2]);
                                    F.lo = ((unsigned long long)(fsl(loc[2], loc[3], 2 * tt) & km[1]) << 32) | (fsl(loc[3], loc[4], 2 * tt) & km[0]);
                                    G.hi = ((unsigned long long)(fsl(rcl[5], rcl[6], 2 * (15 - tt)) & km[3]) << 32) | (fsl(rcl[6], rcl[7], 2 * (15 - tt)) & km[2]);
                                    G.lo = ((unsigned long long)(fsl(rcl[7], rcl[8], 2 * (15 - tt)) & km[1]) << 32) | (fsl(rcl[8], rcl[9], 2 * (15 - tt)) & km[0]);
                                    const key128 cn = key_shr(key_lt(G, F) ? G : F, 128 - 2 * K);
                                    queue_push(wq, warp, lane, ballot, cand[j], cn.hi, cn.lo, db, sink);
                                }
                            }
                        }
                    }
                }
                // slide the register windows
#pragma unroll
                for (int k = 0; k < (int)SEGW - 1; ++k) loc[k] = loc[k + 1];
#pragma unroll
                for (int k = (int)SEGW - 1; k > 0; --k) rcl[k] = rcl[k - 1];
            }
        }
        __syncthreads();                       // every lane is done with this stage's shared memory
        if (tid == 0 && t + 2ull * gridDim.x < ntiles) issue(stage, t + 2ull * gridDim.x);
    }
    queue_drain(wq, warp, lane, db, sink);

    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xFFFFFFFFu, my_valid, o);
    if (lane == 0 && my_valid) atomicAdd(&s_total, my_valid);
    __syncthreads();
    if (tid == 0 && s_total) atomicAdd(a.n_kmers, s_total);
}

// ---------------------------------------------------------------- prep kernels
__device__ __forceinline__ uint32_t ascii_code(unsigned char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
// 8 bases per thread: 2 bytes of packed bases, 1 byte of N mask
__global__ void k_pack_ascii(const unsigned char* text, unsigned long long nbases, unsigned char* bases, unsigned char* nmask,
                             unsigned long long ngroups) {
    unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (t >= ngroups) return;
    uint32_t packed = 0, nm = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        unsigned long long p = t * 8ull + i;
        uint32_t c = p < nbases ? ascii_code(text[p]) : 0u;
        if (c > 3u) { nm |= 0x80u >> i; c = 0; }
        packed |= c << (14 - 2 * i);
    }
    bases[2 * t] = (unsigned char)(packed >> 8);
    bases[2 * t + 1] = (unsigned char)(packed & 0xFFu);
    nmask[t] = (unsigned char)nm;
}
__global__ void k_ascii_to_keys(const unsigned char* text, unsigned long long nslots, uint32_t K, key128* keys, int* bad) {
    unsigned long long s = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    const unsigned char* p = text + s * K;
    key128 k; k.hi = 0; k.lo = 0;
    if (p[0] == 0) { k.hi = ~0ull; k.lo = ~0ull; keys[s] = k; return; }
    for (uint32_t i = 0; i < K; ++i) {
        uint32_t c = ascii_code(p[i]);
        if (c > 3u) { atomicExch(bad, 1); c = 0; }
        k.hi = (k.hi << 2) | (k.lo >> 62);
        k.lo = (k.lo << 2) | c;
    }
    keys[s] = k;
}

constexpr size_t K1_SMEM = sizeof(SharedStage) + sizeof(WarpQueue);

template <int SLOTS, bool HAS_NMASK, bool USE_FILTER>
int launch_probe_k(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    if (db.K == 60) {
        auto kern = k1_decode_canon_probe<SLOTS, HAS_NMASK, 60, USE_FILTER>;
        static bool done[64] = {};      // the attribute is per device
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        kern<<<grid, RT, K1_SMEM, st>>>(a, db);
    } else {
        auto kern = k1_decode_canon_probe<SLOTS, HAS_NMASK, 0, USE_FILTER>;
        static bool done[64] = {};      // the attribute is per device
        int dev = 0;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !done[dev]) {
            CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1_SMEM));
            if (dev >= 0 && dev < 64) done[dev] = true;
        }
        kern<<<grid, RT, K1_SMEM, st>>>(a, db);
    }
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
template <int SLOTS>
int launch_probe_s(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    const bool nm = a.nmask != nullptr, fl = db.nfw != 0;
    if (nm) return fl ? launch_probe_k<SLOTS, true, true>(db, a, st, grid) : launch_probe_k<SLOTS, true, false>(db, a, st, grid);
    return fl ? launch_probe_k<SLOTS, false, true>(db, a, st, grid) : launch_probe_k<SLOTS, false, false>(db, a, st, grid);
}

}  // namespace

int launch_probe(const mlg_ctx* ctx, const DbView& db, const ProbeArgs& a, cudaStream_t st) {
    if (a.r_end <= a.r_begin) return MLG_OK;
    const unsigned long long ntiles = (a.r_end - a.r_begin + RT - 1) / RT;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        ctas_per_sm = 8;   // more CTAs than are resident: the hardware scheduler evens out the tail
        if (const char* s = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(s); if (x >= 1 && x <= 64) ctas_per_sm = x; }
    }
    unsigned long long want = (unsigned long long)ctx->sm_count * ctas_per_sm;
    unsigned grid = (unsigned)(ntiles < want ? ntiles : want);
    if (db.layout == 2) return launch_probe_mz(db, a, st, ctx->sm_count);
    if (db.layout == 1) return launch_probe_sk(db, a, st, ctx->sm_count);
    return db.slots == 8 ? launch_probe_s<8>(db, a, st, grid) : launch_probe_s<4>(db, a, st, grid);
}

int launch_pack_ascii(const unsigned char* text, unsigned long long nbases, unsigned char* bases, unsigned char* nmask,
                      cudaStream_t st) {
    unsigned long long ngroups = (nbases + 7) / 8;
    if (!ngroups) return MLG_OK;
    k_pack_ascii<<<(unsigned)((ngroups + 255) / 256), 256, 0, st>>>(text, nbases, bases, nmask, ngroups);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_ascii_to_keys(const unsigned char* text, unsigned long long nslots, uint32_t K, key128* keys, cudaStream_t st) {
    int* d_bad = nullptr;
    CUDA_TRY(cudaMalloc(&d_bad, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    k_ascii_to_keys<<<(unsigned)((nslots + 255) / 256), 256, 0, st>>>(text, nslots, K, keys, d_bad);
    int bad = 0;
    cudaError_t e = cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_bad);
    if (e != cudaSuccess) { mlg_set_error("ascii_to_keys failed: %s", cudaGetErrorString(e)); return MLG_ERR_CUDA; }
    if (bad) { mlg_set_error("non-ACGT character in a sketch k-mer"); return MLG_ERR_ARG; }
    return MLG_OK;
}
