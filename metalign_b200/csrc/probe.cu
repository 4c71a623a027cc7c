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
#include "mlg_internal.h"

namespace {

#ifndef K1_GROUP
#define K1_GROUP 4                      // windows whose loads are in flight together, per lane
#endif
#ifndef K1_MINCTAS
#define K1_MINCTAS 2                    // __launch_bounds__ minimum CTAs per SM (caps registers at 65536 / (256 * K1_MINCTAS))
#endif
constexpr unsigned RT = 256;            // reads per tile == threads per CTA
constexpr unsigned WARPS = RT / 32;
constexpr unsigned WMAX = 96;           // window starts per segment
constexpr unsigned SEGW = 10;           // 32-bit words of bases per segment (160 bases >= WMAX + 63 - 1)
constexpr unsigned QCAP = 64;           // per-warp queue of exact-path candidates (drained at >= 32)
constexpr unsigned STAGE_B = 16384 + 64;   // staged bytes of packed bases per tile (256 reads x 250 bases fit)
constexpr unsigned STAGE_M = 8192 + 64;    // staged bytes of N mask per tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ unsigned long long bswap64(unsigned long long v) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
// (a:b) << s, upper 32 bits; a is the more significant word; 0 <= s <= 31
__device__ __forceinline__ uint32_t fsl(uint32_t a, uint32_t b, unsigned s) { return __funnelshift_l(b, a, s); }
// reverse the order of the 16 two-bit groups of a word
__device__ __forceinline__ uint32_t rev2_32(uint32_t x) {
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

template <int SLOTS> struct BucketVec;
template <> struct BucketVec<8> { uint32_t w[8]; };
template <> struct BucketVec<4> { uint32_t w[4]; };

// L2 cache policies: the fingerprint table is touched once per probe at random (evict first, do not displace
// anything), the prefilter is the working set that must stay resident (evict last)
__device__ __forceinline__ unsigned long long policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
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

// where the exact path records an occurrence: the counter table, and the list of database k-mers whose
// counter has just reached ci_min (the intersection I, appended exactly once per k-mer)
struct CountSink {
    unsigned char* cnt8;
    uint32_t* present;
    unsigned long long* n_present;      // [0] cursor of present[], [1] cursor of touched[]
    uint32_t* touched;                  // database k-mers seen at least once (the non-zero counters: what a cross-rank exchange needs)
    uint32_t ci_min;
};
// saturating (255) increment of byte counter i via 32-bit CAS
__device__ __forceinline__ void bump_counter(const CountSink& cs, uint32_t i) {
    uint32_t* wp = reinterpret_cast<uint32_t*>(cs.cnt8) + (i >> 2);
    const uint32_t sh = (i & 3u) * 8u;
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    while (((old >> sh) & 0xFFu) < 0xFFu) {
        uint32_t assumed = old;
        old = atomicCAS(wp, assumed, assumed + (1u << sh));
        if (old == assumed) {
            const uint32_t before = (assumed >> sh) & 0xFFu;
            if (before == 0u) cs.touched[atomicAdd(cs.n_present + 1, 1ull)] = i;
            if (before + 1u == cs.ci_min) cs.present[atomicAdd(cs.n_present, 1ull)] = i;
            break;
        }
    }
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
#pragma unroll
            for (int k = 0; k < (int)SEGW; ++k) loc[k] = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) nl[k] = 0;
            if (c) {
                const unsigned long long q = s >> 5;
                const unsigned sh = 2u * (unsigned)(s & 31ull);
                unsigned long long W[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    unsigned long long idx = q + k;
                    if (idx >= a.base_words) idx = a.base_words - 1;
                    W[k] = bswap64(bsrc[idx]);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const unsigned long long v = sh ? ((W[k] << sh) | (W[k + 1] >> (64 - sh))) : W[k];
                    loc[2 * k] = (uint32_t)(v >> 32); loc[2 * k + 1] = (uint32_t)v;
                }
                if (HAS_NMASK) {
                    const unsigned long long qn = s >> 6;
                    const unsigned shn = (unsigned)(s & 63ull);
                    unsigned long long M[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned long long idx = qn + k;
                        if (idx >= a.nmask_words) idx = a.nmask_words - 1;
                        M[k] = bswap64(msrc[idx]);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const unsigned long long v = shn ? ((M[k] << shn) | (M[k + 1] >> (64 - shn))) : M[k];
                        if (2 * k < 5) nl[2 * k] = (uint32_t)(v >> 32);
                        if (2 * k + 1 < 5) nl[2 * k + 1] = (uint32_t)v;
                    }
                }
            }
            // ---- validity of the 96 window starts: i < c and no N in bases [i, i+K)
            uint32_t v0, v1, v2;
            {
                if (HAS_NMASK && __any_sync(0xFFFFFFFFu, (nl[0] | nl[1] | nl[2] | nl[3] | nl[4]) != 0u)) {
                    unsigned cover = 1;
                    while (cover * 2 <= K) {                   // X |= X << cover (towards the smaller base index)
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], cover);
                        nl[4] |= nl[4] << cover;
                        cover *= 2;
                    }
                    const unsigned rest = K - cover;            // < cover <= 32
                    if (rest) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], rest);
                        nl[4] |= nl[4] << rest;
                    }
                }
                // first c bits set (MSB first)
                const uint32_t c0 = c >= 32 ? 0xFFFFFFFFu : (c ? ~(0xFFFFFFFFu >> c) : 0u);
                const uint32_t c1 = c >= 64 ? 0xFFFFFFFFu : (c > 32 ? ~(0xFFFFFFFFu >> (c - 32)) : 0u);
                const uint32_t c2 = c >= 96 ? 0xFFFFFFFFu : (c > 64 ? ~(0xFFFFFFFFu >> (c - 64)) : 0u);
                v0 = ~nl[0] & c0; v1 = ~nl[1] & c1; v2 = ~nl[2] & c2;
            }
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

// ======================================================================================================
// K1, super-k-mer layout (db.layout == 1, K == 60): same lane-per-read walk, but the level-1 bucket PAIR of a
// window is chosen by its MINIMIZER (kmer.cuh) rather than by a hash of the whole k-mer.  Consecutive windows share
// their minimizer for ~18 windows on average, so a lane fetches ~5 bucket pairs (64 bytes each) per 150-base read
// instead of 91 buckets, and the kernel stops being bound by random DRAM sectors.  Every warp is its own pipeline:
// it owns two staging buffers (TMA bulk copies of the packed bases / N mask of its next 32 reads,
// mbarrier-signalled) and never meets the other warps of the CTA at a barrier.  Per block of 16 windows, all of
// it branch-free:
//   phase A  sliding-window minimum of the 45 canonical 16-mers under each window:
//            min(window) = min(suffix of 16-mer block b, whole blocks b+1 [, b+2], prefix of block b+2 / b+3)
//            (van Herk / Gil-Werman on blocks of 16 positions; 16-mer values are recomputed rather than kept:
//            2 funnel shifts + min + multiply-add each).  The window's fingerprint is a mix of its first and last
//            16-mer values, both of which this scan produces anyway.  Where the minimum differs from the one
//            whose pair the lane holds, the new pair is fetched global -> shared by predicated cp.async
//            (up to SK_MAXCH per block; slot 0 is the pair carried in from the previous block);
//   phase B  per window: the half of the held pair that bit 13 of the fingerprint selects is read from shared
//            memory (2 x LDS.128) and its 8 slots are compared with the fingerprint.
// Candidates (fingerprint match, overflowed half, or -- rarely -- a window whose pair found no fetch slot) are
// rebuilt as canonical keys in a rolled loop and go through a per-warp queue to the exact compare.
constexpr unsigned SK_K = 60, SK_W = SK_K - MLG_MIN_M + 1;   // 45 minimizer positions per window
constexpr unsigned SK_MAXCH = 3;                             // new pairs fetched asynchronously per block of 16 windows
constexpr unsigned SK_NSLOT = SK_MAXCH + 1;                  // + the carried one
constexpr uint32_t SK_UNKNOWN = 0xFFFFFFFFu;
static_assert(SK_W == 45, "the block decomposition below is written for 45 positions");
constexpr unsigned WSTAGE_B = 1280 + 64;                     // per-warp staging: 32 reads x 160 bases fit; longer reads are gathered from global
constexpr unsigned WSTAGE_M = 640 + 64;

struct SkStage {
    __align__(16) unsigned char b[WARPS][2][WSTAGE_B];
    __align__(16) unsigned char m[WARPS][2][WSTAGE_M];
};
struct SkSlots {
    uint4 d[SK_NSLOT][4][RT];        // [slot][quarter][thread]: a 64-byte pair per lane; 16-byte accesses of a warp are contiguous
    uint32_t wm[SK_NSLOT + 1][RT];   // minimizer value of the run in each slot (+ one dump entry for runs beyond the last slot)
};
constexpr uint32_t SK_SLOT_STRIDE = 4 * RT * 16, SK_Q_STRIDE = RT * 16, SK_WM_STRIDE = RT * 4;
struct WarpQueueSk {
    unsigned long long hi[WARPS][QCAP];
    unsigned long long lo[WARPS][QCAP];
    uint32_t b[WARPS][QCAP];
    unsigned n[WARPS];
};
// exact compare against the keys of BOTH halves of the pair (their runs of D are adjacent)
__device__ __forceinline__ void probe_exact_sk(const DbView& db, const CountSink& cs, unsigned long long khi, unsigned long long klo,
                                               uint32_t pair) {
    if (pair == SK_UNKNOWN) { key128 c; c.hi = khi; c.lo = klo; pair = (uint32_t)hash_bucket(key_hash_sk(c, db.K, db.bbits), db.bbits) >> 1; }
    uint32_t s = db.bstart[2ull * pair], e = db.bstart[2ull * pair + 2];
    for (uint32_t i = s; i < e; ++i) {
        key128 d = db.D_key[i];
        if (d.hi == khi && d.lo == klo) { bump_counter(cs, i); return; }
    }
}
__device__ __noinline__ void queue_drain_sk(WarpQueueSk& q, unsigned warp, unsigned lane, const DbView& db, const CountSink& cs) {
    __syncwarp();
    const unsigned n = q.n[warp];
    for (unsigned i = lane; i < n; i += 32) probe_exact_sk(db, cs, q.hi[warp][i], q.lo[warp][i], q.b[warp][i]);
    __syncwarp();
    if (lane == 0) q.n[warp] = 0;
    __syncwarp();
}
__device__ __forceinline__ void queue_push_sk(WarpQueueSk& q, unsigned warp, unsigned lane, unsigned ballot, bool cand,
                                              unsigned long long khi, unsigned long long klo, uint32_t pair, const DbView& db,
                                              const CountSink& cs) {
    const unsigned base = q.n[warp];
    if (cand) {
        const unsigned i = base + __popc(ballot & ((1u << lane) - 1u));
        q.hi[warp][i] = khi; q.lo[warp][i] = klo; q.b[warp][i] = pair;
    }
    __syncwarp();
    const unsigned total = base + __popc(ballot);
    if (lane == 0) q.n[warp] = total;
    __syncwarp();
    if (total >= 32) queue_drain_sk(q, warp, lane, db, cs);
}
// if (p): fetch the 64-byte bucket pair at src into the lane's slot (four 16-byte cp.async)
__device__ __forceinline__ void sk_fetch_if(bool p, uint32_t slot_addr, const uint32_t* src) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %0, 0;\n"
        "@q cp.async.cg.shared.global [%1], [%2], 16;\n"
        "@q cp.async.cg.shared.global [%1+4096], [%2+16], 16;\n"
        "@q cp.async.cg.shared.global [%1+8192], [%2+32], 16;\n"
        "@q cp.async.cg.shared.global [%1+12288], [%2+48], 16;\n"
        "}\n" ::"r"((uint32_t)p), "r"(slot_addr), "l"(src)
        : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
static_assert(SK_Q_STRIDE == 4096, "sk_fetch_if hard-codes the quarter stride");
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// mixed value of the canonical 16-mer at position 16*j + i of the current block (0 <= j <= 3): its forward strand is
// bases [16j+i, 16j+i+16) of loc[], its reverse complement starts at base 139 - (16j+i) of rcl[] (the complement of
// base x of the block sits at base 154 - x of rcl[], see the alignment of rcl[] in the kernel)
__device__ __forceinline__ uint32_t sk_mmer(const uint32_t (&loc)[SEGW], const uint32_t (&rcl)[SEGW], int j, int i) {
    const uint32_t f = fsl(loc[j], loc[j + 1], 2 * i);
    const int a = i <= 11 ? 8 - j : 7 - j, o = i <= 11 ? 11 - i : 27 - i;
    const uint32_t r = fsl(rcl[a], rcl[a + 1], 2 * o);
    return mmer_mix(f, r);
}

template <bool HAS_NMASK>
__global__ void __launch_bounds__(RT, K1_MINCTAS) k1_superkmer_probe(ProbeArgs a, DbView db) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SkStage& stg = *reinterpret_cast<SkStage*>(smem_raw);
    SkSlots& slots = *reinterpret_cast<SkSlots*>(smem_raw + sizeof(SkStage));
    WarpQueueSk& wq = *reinterpret_cast<WarpQueueSk*>(smem_raw + sizeof(SkStage) + sizeof(SkSlots));
    __shared__ __align__(8) unsigned long long mbar[WARPS][2];
    __shared__ unsigned long long s_bw0[WARPS][2], s_mw0[WARPS][2];
    __shared__ unsigned s_staged[WARPS][2];

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    constexpr unsigned K = SK_K;
    const unsigned long long nreads = a.r_end - a.r_begin;
    const unsigned long long ntiles = (nreads + 31) / 32;                 // a tile = the 32 reads of one warp pass
    // tiles are handed out dynamically (one global atomic per 32 reads): SMs differ in how fast they get through
    // their tiles (die, L2 distance), and a static split leaves the fast ones idle at the end
    auto next_tile = [&]() -> unsigned long long {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.tile_counter, 1ull);
        return __shfl_sync(0xFFFFFFFFu, t, 0);
    };
    const unsigned qshift = 33u - db.bbits;      // pair index = top bbits-1 bits of the minimizer's bucket hash; 2 <= bbits <= 31
    const unsigned long long pol_stream = policy_evict_first();
    const CountSink sink{a.cnt8, a.present, a.n_present, a.touched, a.ci_min};
    const uint32_t slot_base = smem_u32(&slots.d[0][0][tid]), wm_base = smem_u32(&slots.wm[0][tid]);
    const uint32_t wm_dump = wm_base + SK_NSLOT * SK_WM_STRIDE;

    if (lane == 0) {
        mbar_init(&mbar[warp][0], 1);
        mbar_init(&mbar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        wq.n[warp] = 0;
    }
    __syncthreads();

    // producer side (lane 0 of each warp): stage the stream range of one warp tile, if it fits
    auto issue = [&](unsigned stage, unsigned long long t) {
        const unsigned long long r0 = a.r_begin + t * 32ull;
        const unsigned long long r1 = (r0 + 32 < a.r_end) ? r0 + 32 : a.r_end;
        const unsigned long long p0 = a.off ? a.off[r0] : r0 * (unsigned long long)a.read_len;
        const unsigned long long p1 = a.off ? a.off[r1] : r1 * (unsigned long long)a.read_len;
        unsigned long long bw0 = (p0 >> 5) & ~1ull;
        unsigned long long bw1 = ((p1 + 31) >> 5) + 6;
        if (bw1 > a.base_words) bw1 = a.base_words;
        bw1 = (bw1 + 1) & ~1ull;
        unsigned long long mw0 = (p0 >> 6) & ~1ull;
        unsigned long long mw1 = ((p1 + 63) >> 6) + 4;
        if (HAS_NMASK) { if (mw1 > a.nmask_words) mw1 = a.nmask_words; mw1 = (mw1 + 1) & ~1ull; }
        const unsigned long long bytes_b = (bw1 - bw0) * 8ull, bytes_m = HAS_NMASK ? (mw1 - mw0) * 8ull : 0ull;
        const bool fits = bw1 > bw0 && bytes_b <= WSTAGE_B && bytes_m <= WSTAGE_M;
        s_bw0[warp][stage] = bw0; s_mw0[warp][stage] = mw0; s_staged[warp][stage] = fits ? 1u : 0u;
        if (fits) {
            mbar_expect_tx(&mbar[warp][stage], (uint32_t)(bytes_b + bytes_m));
            bulk_g2s(&stg.b[warp][stage][0], a.bases + bw0, (uint32_t)bytes_b, &mbar[warp][stage], pol_stream);
            if (HAS_NMASK && bytes_m) bulk_g2s(&stg.m[warp][stage][0], a.nmask + mw0, (uint32_t)bytes_m, &mbar[warp][stage], pol_stream);
        } else {
            mbar_arrive(&mbar[warp][stage]);
        }
    };

    constexpr uint32_t KM0 = 0xFFFFFFFFu << (128 - 2 * K);     // K = 60: the low word of a top-aligned k-mer keeps 24 bits
    unsigned long long my_valid = 0;
    unsigned my_fetch = 0;
    // the minimizer value whose pair this lane holds in slot 0 (have: false until the first block has been processed)
    uint32_t held_wm = 0;
    bool have = false;

    unsigned it = 0;
    unsigned long long t = next_tile(), t_ahead = next_tile();       // the tile being processed and the one staged behind it
    if (lane == 0) {
        if (t < ntiles) issue(0, t);
        if (t_ahead < ntiles) issue(1, t_ahead);
    }
    for (; t < ntiles; ++it) {
        const unsigned stage = it & 1u, parity = (it >> 1) & 1u;
        __syncwarp();
        mbar_wait(&mbar[warp][stage], parity);

        const unsigned long long r = a.r_begin + t * 32ull + lane;
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
        const bool staged = s_staged[warp][stage] != 0;
        const unsigned long long* bsrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.b[warp][stage][0]) - s_bw0[warp][stage] : a.bases;
        const unsigned long long* msrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.m[warp][stage][0]) - s_mw0[warp][stage] : a.nmask;

        for (unsigned seg = 0; seg < max_seg; ++seg) {
            const unsigned c = seg < nseg ? (unsigned)((nw - (unsigned long long)seg * WMAX) < WMAX ? (nw - (unsigned long long)seg * WMAX) : WMAX) : 0u;
            const unsigned long long s = R0 + (unsigned long long)seg * WMAX;

            uint32_t loc[SEGW];
            uint32_t nl[5];
#pragma unroll
            for (int k = 0; k < (int)SEGW; ++k) loc[k] = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) nl[k] = 0;
            if (c) {
                const unsigned long long q = s >> 5;
                const unsigned sh = 2u * (unsigned)(s & 31ull);
                unsigned long long W[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    unsigned long long idx = q + k;
                    if (idx >= a.base_words) idx = a.base_words - 1;
                    W[k] = bswap64(bsrc[idx]);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const unsigned long long v = sh ? ((W[k] << sh) | (W[k + 1] >> (64 - sh))) : W[k];
                    loc[2 * k] = (uint32_t)(v >> 32); loc[2 * k + 1] = (uint32_t)v;
                }
                if (HAS_NMASK) {
                    const unsigned long long qn = s >> 6;
                    const unsigned shn = (unsigned)(s & 63ull);
                    unsigned long long M[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned long long idx = qn + k;
                        if (idx >= a.nmask_words) idx = a.nmask_words - 1;
                        M[k] = bswap64(msrc[idx]);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const unsigned long long v = shn ? ((M[k] << shn) | (M[k + 1] >> (64 - shn))) : M[k];
                        if (2 * k < 5) nl[2 * k] = (uint32_t)(v >> 32);
                        if (2 * k + 1 < 5) nl[2 * k + 1] = (uint32_t)v;
                    }
                }
            }
            uint32_t v0, v1, v2;
            {
                if (HAS_NMASK && __any_sync(0xFFFFFFFFu, (nl[0] | nl[1] | nl[2] | nl[3] | nl[4]) != 0u)) {
                    unsigned cover = 1;
                    while (cover * 2 <= K) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], cover);
                        nl[4] |= nl[4] << cover;
                        cover *= 2;
                    }
                    const unsigned rest = K - cover;
                    if (rest) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], rest);
                        nl[4] |= nl[4] << rest;
                    }
                }
                const uint32_t c0 = c >= 32 ? 0xFFFFFFFFu : (c ? ~(0xFFFFFFFFu >> c) : 0u);
                const uint32_t c1 = c >= 64 ? 0xFFFFFFFFu : (c > 32 ? ~(0xFFFFFFFFu >> (c - 32)) : 0u);
                const uint32_t c2 = c >= 96 ? 0xFFFFFFFFu : (c > 64 ? ~(0xFFFFFFFFu >> (c - 64)) : 0u);
                v0 = ~nl[0] & c0; v1 = ~nl[1] & c1; v2 = ~nl[2] & c2;
            }
            my_valid += __popc(v0) + __popc(v1) + __popc(v2);
            if (__all_sync(0xFFFFFFFFu, (v0 | v1 | v2) == 0u)) continue;

            // reverse complement of the segment: the complement of base x sits at base 154 - x of rcl[]
            uint32_t rcl[SEGW];
            {
                uint32_t t160[SEGW + 1];
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) t160[k] = rev2_32(~loc[SEGW - 1 - k]);
                t160[SEGW] = 0;
                constexpr unsigned bs = 2u * (160u - (WMAX + K - 1u));       // 10 bits dropped at the front
                static_assert(bs < 32, "alignment shift must stay inside one word");
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) rcl[k] = fsl(t160[k], t160[k + 1], bs);
            }

            // minimizer state carried from block to block: A1 = min of 16-mer block b+1, P = prefix of block b+2 up to index 11
            uint32_t A1 = 0xFFFFFFFFu, P = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 16; ++i) A1 = min(A1, sk_mmer(loc, rcl, 1, i));
#pragma unroll
            for (int i = 0; i < 12; ++i) P = min(P, sk_mmer(loc, rcl, 2, i));

#pragma unroll 1
            for (int blk = 0; blk < (int)(WMAX / 16); ++blk) {
                if (__all_sync(0xFFFFFFFFu, (v0 | v1 | v2) == 0u)) break;       // nothing valid from here to the end of the segment
                const uint32_t vb = v0 & 0xFFFF0000u;
                v0 = fsl(v0, v1, 16); v1 = fsl(v1, v2, 16); v2 <<= 16;

                // ---- phase A: window minima, fingerprints, runs of equal minimizer (validity is ignored here: a window
                //      that is not valid costs at most a wasted fetch; it is masked out of the candidates below)
                uint32_t fpv[16];                 // value of the window's first 16-mer, then the window's fingerprint
                uint32_t chgraw = 0, wm_first = 0;
                {
                    uint32_t Suf[16];
                    {
                        uint32_t sm = 0xFFFFFFFFu;
#pragma unroll
                        for (int i = 15; i >= 0; --i) { fpv[i] = sk_mmer(loc, rcl, 0, i); sm = min(sm, fpv[i]); Suf[i] = sm; }
                    }
                    uint32_t base12 = A1, A2 = 0, pw = 0;
                    uint32_t lst = wm_base + SK_WM_STRIDE;          // where the next run's minimizer goes
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) {
                        uint32_t wm, he;
                        if (tt < 4) {
                            he = sk_mmer(loc, rcl, 2, 12 + tt);
                            P = min(P, he);
                            wm = min(min(Suf[tt], A1), P);
                            if (tt == 3) { A2 = P; base12 = min(A1, A2); P = 0xFFFFFFFFu; }
                        } else {
                            he = sk_mmer(loc, rcl, 3, tt - 4);
                            P = min(P, he);
                            wm = min(min(Suf[tt], base12), P);
                        }
                        fpv[tt] += he;                              // mixed into the fingerprint after the fetches are issued
                        if (tt == 0) { wm_first = wm; sts32(wm_base, wm); }
                        else if (wm != pw) {                       // a new run starts at window tt
                            sts32(min(lst, wm_dump), wm);
                            lst += SK_WM_STRIDE;
                            chgraw |= 1u << tt;
                        }
                        pw = wm;
                    }
                    A1 = A2;
                }
                // runs 0..3 of the block live in slots 0..3; windows of later runs (rare) take the exact path with an unknown pair
                uint32_t chg, ovf;
                {
                    uint32_t t3 = chgraw;
                    t3 &= t3 - 1u; t3 &= t3 - 1u; t3 &= t3 - 1u;      // changes beyond the third
                    chg = chgraw ^ t3;
                    ovf = t3 ? ~((t3 & (0u - t3)) - 1u) & 0xFFFFu : 0u;   // every window from the fourth change on
                }
                const unsigned nrun = 1u + __popc(chg);
                // run 0 continues the pair carried in slot 0 unless its minimizer differs; runs 1.. are always new
                {
                    const bool need0 = !have || wm_first != held_wm;
                    sk_fetch_if(need0, slot_base, db.T1 + (unsigned long long)((wm_first * MLG_BKT_MULT) >> qshift) * 16ull);
                    my_fetch += need0 ? 1u : 0u;
#pragma unroll
                    for (unsigned rr = 1; rr < SK_NSLOT; ++rr) {
                        const uint32_t w = lds32(wm_base + rr * SK_WM_STRIDE);
                        sk_fetch_if(rr < nrun, slot_base + rr * SK_SLOT_STRIDE, db.T1 + (unsigned long long)((w * MLG_BKT_MULT) >> qshift) * 16ull);
                    }
                    my_fetch += nrun - 1u;
                }
                // fingerprints from the sums (kmer.cuh: sk_fp) while the fetches are in flight
#pragma unroll
                for (int tt = 0; tt < 16; ++tt) fpv[tt] = sk_fp(fpv[tt], 0u);
                cp_async_wait_all();

                // ---- phase B: fingerprints against the half of the held pair that the fingerprint selects
                if (!__all_sync(0xFFFFFFFFu, vb == 0u)) {
                    uint32_t rd = slot_base;
                    uint32_t candm = ovf;             // windows that must take the exact path
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) {
                        rd += ((chg >> tt) & 1u) * SK_SLOT_STRIDE;
                        const uint32_t fp = fpv[tt];
                        const uint32_t ha = rd + (fp & (2u * SK_Q_STRIDE));              // bit 13 of the fingerprint: which half
                        const uint4 x = lds128(ha), y = lds128(ha + SK_Q_STRIDE);
                        // a match, or a half with no free slot left (it may have overflowed: exact path)
                        const bool hit = (x.x == fp) | (x.y == fp) | (x.z == fp) | (x.w == fp) | (y.x == fp) | (y.y == fp) | (y.z == fp) | (y.w != 0u);
                        if (hit) candm |= 1u << tt;
                    }
                    candm &= __brev(vb);              // bit tt of brev(vb) = validity of window tt
                    if (__any_sync(0xFFFFFFFFu, candm != 0u)) {
                        // rare path, rolled: rebuild the canonical key of every candidate window and queue it
#pragma unroll 1
                        for (unsigned tt = 0; tt < 16; ++tt) {
                            const bool cnd = (candm >> tt) & 1u;
                            const unsigned ballot = __ballot_sync(0xFFFFFFFFu, cnd);
                            if (!ballot) continue;
                            const unsigned sf = 2u * tt, sr = 30u - sf;
                            key128 F, G;
                            F.hi = ((unsigned long long)fsl(loc[0], loc[1], sf) << 32) | fsl(loc[1], loc[2], sf);
                            F.lo = ((unsigned long long)fsl(loc[2], loc[3], sf) << 32) | (fsl(loc[3], loc[4], sf) & KM0);
                            G.hi = ((unsigned long long)fsl(rcl[5], rcl[6], sr) << 32) | fsl(rcl[6], rcl[7], sr);
                            G.lo = ((unsigned long long)fsl(rcl[7], rcl[8], sr) << 32) | (fsl(rcl[8], rcl[9], sr) & KM0);
                            const key128 cn = key_shr(key_lt(G, F) ? G : F, 128 - 2 * K);
                            // the pair the lane held at window tt: slot = number of changes at or before tt
                            uint32_t pr = SK_UNKNOWN;
                            if (cnd && !((ovf >> tt) & 1u)) pr = (slots.wm[__popc(chg & ((2u << tt) - 1u))][tid] * MLG_BKT_MULT) >> qshift;
                            queue_push_sk(wq, warp, lane, ballot, cnd, cn.hi, cn.lo, pr, db, sink);
                        }
                    }
                }
                // the pair of the last run in a slot moves to slot 0 for the next block
                have = true;
                held_wm = lds32(wm_base + (nrun - 1u) * SK_WM_STRIDE);
                if (nrun > 1u) {
                    const uint32_t from = slot_base + (nrun - 1u) * SK_SLOT_STRIDE;
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) sts128(slot_base + qd * SK_Q_STRIDE, lds128(from + qd * SK_Q_STRIDE));
                }
                // slide the register windows by one word
#pragma unroll
                for (int k = 0; k < (int)SEGW - 1; ++k) loc[k] = loc[k + 1];
#pragma unroll
                for (int k = (int)SEGW - 1; k > 0; --k) rcl[k] = rcl[k - 1];
            }
        }
        __syncwarp();
        // this stage is free again: stage the tile after next into it
        const unsigned long long t_new = next_tile();
        if (lane == 0 && t_new < ntiles) issue(stage, t_new);
        t = t_ahead; t_ahead = t_new;
    }
    queue_drain_sk(wq, warp, lane, db, sink);

    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xFFFFFFFFu, my_valid, o);
    my_fetch = __reduce_add_sync(0xFFFFFFFFu, my_fetch);
    // every warp reports for itself: no CTA-wide barrier at the end either
    if (lane == 0 && my_valid) atomicAdd(a.n_kmers, my_valid);
    if (lane == 0 && my_fetch) atomicAdd(a.n_kmers + 1, (unsigned long long)my_fetch);
}
constexpr size_t K1SK_SMEM = sizeof(SkStage) + sizeof(SkSlots) + sizeof(WarpQueueSk);

template <bool HAS_NMASK>
int launch_probe_sk(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    auto kern = k1_superkmer_probe<HAS_NMASK>;
    static bool done[64] = {};          // the attribute is per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1SK_SMEM));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    CUDA_TRY(cudaMemsetAsync(a.tile_counter, 0, 8, st));
    kern<<<grid, RT, K1SK_SMEM, st>>>(a, db);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}

// ======================================================================================================
// K1, minimizer-bitmap layout (db.layout == 2, K == 60; definitions in kmer.cuh).  Same lane-per-read walk and
// per-warp TMA pipelines as the super-k-mer kernel above, but the minimizer is a 32-mer whose (order << 6 | position)
// value is two multiply-adds, level 1 is a bit array over minimizer IDENTITIES, and there is no per-window compare:
//   phase A  sliding minimum over the 29 32-mers under each window (van Herk / Gil-Werman on blocks of 16 positions)
//            and the runs of equal minimizer; values of new runs go to a per-lane list in shared memory;
//   lookup   for up to MZ_MAXRUN runs per block of 16 windows the lane fetches the two halves of the minimizer from its
//            copy of the segment in shared memory (the position is in the value), mixes their identity and loads the
//            bit-array word (predicated LDG);
//   deferred the words are looked at one block LATER (after the next block's phase A, which hides the DRAM latency):
//            a run whose bit is set becomes an ITEM (source lane, first window of the block, identity, 16-bit mask of
//            its valid windows) in a per-warp ring in shared memory.
// Items are rare (bit density <= 1/32, plus the true hits); the ring is drained 32 at a time, one item per lane: the
// lane rebuilds the canonical key of each window of the item from the staged bases and compares it with the database
// k-mers filed under that identity (bucket index on the identity's high word, plus the alias table).
#ifndef MZ_MINCTAS
#define MZ_MINCTAS 2
#endif
constexpr unsigned MZ_MAXRUN = 4;             // runs per block whose bit-array word is fetched ahead (the carried one + 3 new)
constexpr unsigned MZ_QDRAIN = 32;            // per-warp item list: drained whenever this many are waiting ...
constexpr unsigned MZ_QCAP = MZ_QDRAIN + 128; // ... and one block of 16 windows adds at most 4 x 32
constexpr unsigned MZ_LIST = 16;              // run list rows: a block of 16 windows starts at most 15 new runs (row 0 unused)
constexpr uint32_t MZ_M64 = MLG_MZ_ORD_MULT << 6;   // the multiplier carries the << 6 of (order << 6 | position)
struct MzShared {
    SkStage stg;
    uint32_t wm[MZ_LIST][RT];                 // [r][thread]: value of the r-th run START of the current block (r >= 1)
    uint32_t seq[SEGW + 1][RT];               // [word][thread]: the lane's current segment (160 bases, top-aligned words)
    uint32_t qa[WARPS][MZ_QCAP];              // item: source lane | index of the block's first window in the read << 5
    uint32_t qb[WARPS][MZ_QCAP];              // item: windows of the block to compare exactly (bit tt = window tt) | base of the
                                              //       minimizer relative to the block's first window << 16
    uint32_t qoff[WARPS][32];                 // drain: windows before each item of the batch
    uint32_t qn[WARPS];                       // items waiting
    unsigned long long r0s[WARPS][32];        // stream position (in bases) of each lane's read of the current tile
};
constexpr uint32_t MZ_ROW = RT * 4;           // byte stride between rows of wm[] / seq[]

__device__ __forceinline__ uint32_t ldg_bitmap_if(bool p, const uint32_t* ptr) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %1, 0;\n"
        "mov.u32 %0, 0;\n"
        "@q ld.global.nc.L1::no_allocate.b32 %0, [%2];\n"
        "}\n" : "=r"(r) : "r"((uint32_t)p), "l"(ptr));
    return r;
}
// (order << 6 | position) of the 32-mer at position 16j+i of the current block: first half = bases [16j+i, +16) of
// loc[], reverse complement of the second half = the 16-mer at 16(j+1)+i seen through rcl[] (see sk_mmer)
__device__ __forceinline__ uint32_t mz_val(const uint32_t (&loc)[SEGW], const uint32_t (&rcl)[SEGW], int j, int i) {
    const uint32_t a = fsl(loc[j], loc[j + 1], 2 * i);
    const int ra = i <= 11 ? 7 - j : 6 - j, ro = i <= 11 ? 11 - i : 27 - i;
    const uint32_t b = fsl(rcl[ra], rcl[ra + 1], 2 * ro);
    return b * MZ_M64 + (a * MZ_M64 + (uint32_t)(16 * j + i));
}
// 64 bases of the packed stream starting at base p, top-aligned (hi = the first 32)
__device__ __forceinline__ void mz_bases64(const unsigned long long* bsrc, unsigned long long base_words, unsigned long long p,
                                           unsigned long long& hi, unsigned long long& lo) {
    const unsigned long long q = p >> 5;
    const unsigned sh = 2u * (unsigned)(p & 31ull);
    unsigned long long W[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long idx = q + k;
        if (idx >= base_words) idx = base_words - 1;
        W[k] = bswap64(bsrc[idx]);
    }
    hi = sh ? ((W[0] << sh) | (W[1] >> (64 - sh))) : W[0];
    lo = sh ? ((W[1] << sh) | (W[2] >> (64 - sh))) : W[1];
}
// Exact compare of ONE window of an item (one lane), in two stages.  Stage 1: the window starts at base p, its minimizer
// at base pm of the stream; identity, bucket range and the first two database k-mers of the bucket (both loads in flight
// together).  (Two windows per lane and round were tried: the drain's code doubles and the kernel starts missing in the
// instruction cache, which costs more than the extra memory parallelism gains.)
struct MzWin {
    key128 cn, d0, d1;           // canonical key of the window; first two k-mers of the bucket
    uint32_t s, e, j0, j1;       // bucket range in D, alias range
};
__device__ __forceinline__ void mz_window_begin(MzWin& w, bool on, unsigned long long p, unsigned long long pm, bool untested,
                                                const unsigned long long* bsrc, unsigned long long base_words, const DbView& db) {
    w.s = w.e = w.j0 = w.j1 = 0;
    w.cn.hi = w.cn.lo = 0; w.d0 = w.cn; w.d1 = w.cn;
    if (!on) return;
    unsigned long long mh, ml;
    mz_bases64(bsrc, base_words, pm, mh, ml);
    const uint32_t ha = (uint32_t)(mh >> 32), hb = rev2_32(~(uint32_t)mh);
    const uint32_t zhi = mz_ident_hi(ha, hb), zlo = mz_ident_lo(ha, hb);
    if (untested) {               // a run beyond the fourth of its block: its level-1 word was not fetched ahead
        const unsigned long long idx = mz_bit_index(((unsigned long long)zhi << 32) | zlo, db.fbits);
        const uint32_t f = db.F[idx >> 5];
        if (!((f >> (zlo & 31u)) & (f >> mz_bit2(zhi)) & 1u)) return;
    }
    const uint32_t bucket = zhi >> (32u - db.bbits);
    w.s = db.bstart[bucket]; w.e = db.bstart[bucket + 1];
    // K-mers filed under a second identity (order ties) are rare: a 2^16-bit array says whether to look at all
    if (db.n_alias && ((db.alias_bloom[(zlo & 0xFFFFu) >> 5] >> (zlo & 31u)) & 1u)) {
        const unsigned long long z = ((unsigned long long)zhi << 32) | zlo;
        uint32_t lo = 0, hi = db.n_alias;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (db.alias_z[mid] < z) lo = mid + 1; else hi = mid; }
        w.j0 = w.j1 = lo;
        while (w.j1 < db.n_alias && db.alias_z[w.j1] == z) ++w.j1;
    }
    key128 F;                                                       // 64 bases from p, top-aligned
    mz_bases64(bsrc, base_words, p, F.hi, F.lo);
    F = key_shr(F, 128 - 2 * SK_K);                                 // the 60-mer, bottom-aligned
    const key128 G = key_rc(F, SK_K);
    w.cn = key_lt(G, F) ? G : F;
    if (w.s < w.e) w.d0 = db.D_key[w.s];
    if (w.s + 1u < w.e) w.d1 = db.D_key[w.s + 1u];
}
// stage 2: compare with the database k-mers filed under the identity, count a match
__device__ __forceinline__ void mz_window_end(const MzWin& w, const DbView& db, const CountSink& cs) {
    if (w.s < w.e && w.d0.hi == w.cn.hi && w.d0.lo == w.cn.lo) { bump_counter(cs, w.s); return; }
    if (w.s + 1u < w.e && w.d1.hi == w.cn.hi && w.d1.lo == w.cn.lo) { bump_counter(cs, w.s + 1u); return; }
    for (uint32_t i = w.s + 2u; i < w.e; ++i) {
        const key128 d = db.D_key[i];
        if (d.hi == w.cn.hi && d.lo == w.cn.lo) { bump_counter(cs, i); return; }
    }
    for (uint32_t j = w.j0; j < w.j1; ++j) {
        const uint32_t i = db.alias_i[j];
        const key128 d = db.D_key[i];
        if (d.hi == w.cn.hi && d.lo == w.cn.lo) { bump_counter(cs, i); return; }
    }
}
// exact compare of every window of every waiting item of one warp, one WINDOW per lane (items have 1..16 windows)
__device__ __noinline__ void mz_drain(uint32_t* qn, const uint32_t* qa, const uint32_t* qb, uint32_t* qoff, const unsigned long long* r0s,
                                      const unsigned long long* bsrc, unsigned long long base_words, const DbView& db, const CountSink& cs) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp();
    const uint32_t n = *qn;
    for (uint32_t base = 0; base < n; base += 32u) {
        const uint32_t i = base + lane;
        const uint32_t cnt = i < n ? (uint32_t)__popc(qb[i] & 0xFFFFu) : 0u;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, inc, o); if ((int)lane >= o) inc += u; }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        qoff[lane] = inc - cnt;
        __syncwarp();
        for (uint32_t w = lane; w < total; w += 32u) {
            uint32_t j = 0;                              // the last item of the batch that starts at or before window w
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1) if (qoff[j + step] <= w) j += step;
            const uint32_t ia = qa[base + j], kb = qb[base + j];
            const uint32_t tt = __fns(kb & 0xFFFFu, 0u, (int)(w - qoff[j]) + 1);
            const unsigned long long pb = r0s[ia & 31u] + (ia >> 5);
            MzWin win;
            mz_window_begin(win, true, pb + tt, pb + ((kb >> 16) & 63u), (kb >> 31) != 0u, bsrc, base_words, db);
            mz_window_end(win, db, cs);
        }
        __syncwarp();
    }
    if (lane == 0) *qn = 0;
    __syncwarp();
}

template <bool HAS_NMASK>
__global__ void __launch_bounds__(RT, MZ_MINCTAS) k1_minimizer_probe(ProbeArgs a, DbView db) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MzShared& sm = *reinterpret_cast<MzShared*>(smem_raw);
    SkStage& stg = sm.stg;
    __shared__ __align__(8) unsigned long long mbar[WARPS][2];
    __shared__ unsigned long long s_bw0[WARPS][2], s_mw0[WARPS][2];
    __shared__ unsigned s_staged[WARPS][2];

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    constexpr unsigned K = SK_K;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned long long nreads = a.r_end - a.r_begin;
    const unsigned long long ntiles = (nreads + 31) / 32;                 // a tile = the 32 reads of one warp pass
    auto next_tile = [&]() -> unsigned long long {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(a.tile_counter, 1ull);
        return __shfl_sync(FULL, t, 0);
    };
    const unsigned long long pol_stream = policy_evict_first();
    const CountSink sink{a.cnt8, a.present, a.n_present, a.touched, a.ci_min};
    const uint32_t wm_base = smem_u32(&sm.wm[0][tid]), seq_base = smem_u32(&sm.seq[0][tid]);
    const uint32_t* const MB = db.F;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // bit index = identity & (2^fbits - 1): mask of its low word, and of its high word (0 up to 2^32 bits)
    const uint32_t fmask_lo = db.fbits >= 32u ? 0xFFFFFFFFu : ((1u << db.fbits) - 1u);
    const uint32_t fmask_hi = db.fbits > 32u ? ((1u << (db.fbits - 32u)) - 1u) : 0u;

    if (lane == 0) {
        mbar_init(&mbar[warp][0], 1);
        mbar_init(&mbar[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](unsigned stage, unsigned long long t) {
        const unsigned long long r0 = a.r_begin + t * 32ull;
        const unsigned long long r1 = (r0 + 32 < a.r_end) ? r0 + 32 : a.r_end;
        const unsigned long long p0 = a.off ? a.off[r0] : r0 * (unsigned long long)a.read_len;
        const unsigned long long p1 = a.off ? a.off[r1] : r1 * (unsigned long long)a.read_len;
        unsigned long long bw0 = (p0 >> 5) & ~1ull;
        unsigned long long bw1 = ((p1 + 31) >> 5) + 6;
        if (bw1 > a.base_words) bw1 = a.base_words;
        bw1 = (bw1 + 1) & ~1ull;
        unsigned long long mw0 = (p0 >> 6) & ~1ull;
        unsigned long long mw1 = ((p1 + 63) >> 6) + 4;
        if (HAS_NMASK) { if (mw1 > a.nmask_words) mw1 = a.nmask_words; mw1 = (mw1 + 1) & ~1ull; }
        const unsigned long long bytes_b = (bw1 - bw0) * 8ull, bytes_m = HAS_NMASK ? (mw1 - mw0) * 8ull : 0ull;
        const bool fits = bw1 > bw0 && bytes_b <= WSTAGE_B && bytes_m <= WSTAGE_M;
        s_bw0[warp][stage] = bw0; s_mw0[warp][stage] = mw0; s_staged[warp][stage] = fits ? 1u : 0u;
        if (fits) {
            mbar_expect_tx(&mbar[warp][stage], (uint32_t)(bytes_b + bytes_m));
            bulk_g2s(&stg.b[warp][stage][0], a.bases + bw0, (uint32_t)bytes_b, &mbar[warp][stage], pol_stream);
            if (HAS_NMASK && bytes_m) bulk_g2s(&stg.m[warp][stage][0], a.nmask + mw0, (uint32_t)bytes_m, &mbar[warp][stage], pol_stream);
        } else {
            mbar_arrive(&mbar[warp][stage]);
        }
    };

    unsigned long long my_valid = 0;
    unsigned my_fetch = 0;
    const unsigned long long* bsrc = a.bases;
    if (lane == 0) sm.qn[warp] = 0;
    __syncwarp();

    // the two halves (a, b) of the 32-mer at base P of the lane's segment copy (b = reverse complement of the second half)
    auto halves_at = [&](uint32_t P, uint32_t& ha, uint32_t& hb) {
        const uint32_t ad = seq_base + (P >> 4) * MZ_ROW;
        const uint32_t w0 = lds32(ad), w1 = lds32(ad + MZ_ROW), w2 = lds32(ad + 2u * MZ_ROW);
        const unsigned sh = 2u * (P & 15u);
        ha = fsl(w0, w1, sh);
        hb = rev2_32(~fsl(w1, w2, sh));
    };
    auto drain = [&]() {
        mz_drain(&sm.qn[warp], &sm.qa[warp][0], &sm.qb[warp][0], &sm.qoff[warp][0], &sm.r0s[warp][0], bsrc, a.base_words, db, sink);
    };
    // lanes with p append one item: windows `ik` of the block whose first window is `ia`, minimizer at base `rel` of the block
    auto push = [&](bool p, uint32_t ia, uint32_t ik, uint32_t rel) {
        if (p) {
            const uint32_t i = atomicAdd(&sm.qn[warp], 1u);
            sm.qa[warp][i] = ia; sm.qb[warp][i] = ik | (rel << 16);
        }
    };
    auto drain_if_full = [&]() {
        __syncwarp();
        if (sm.qn[warp] >= MZ_QDRAIN) drain();
    };

    unsigned it = 0;
    unsigned long long t = next_tile(), t_ahead = next_tile();       // the tile being processed and the one staged behind it
    if (lane == 0) {
        if (t < ntiles) issue(0, t);
        if (t_ahead < ntiles) issue(1, t_ahead);
    }
    for (; t < ntiles; ++it) {
        const unsigned stage = it & 1u, parity = (it >> 1) & 1u;
        __syncwarp();
        mbar_wait(&mbar[warp][stage], parity);

        const unsigned long long r = a.r_begin + t * 32ull + lane;
        const bool active = r < a.r_end;
        unsigned long long R0 = 0, R1 = 0;
        if (active) {
            R0 = a.off ? a.off[r] : r * (unsigned long long)a.read_len;
            R1 = a.off ? a.off[r + 1] : R0 + a.read_len;
        }
        sm.r0s[warp][lane] = R0;
        const unsigned long long len = R1 - R0;
        const unsigned long long nw = len >= K ? len - K + 1 : 0ull;
        const unsigned nseg = (unsigned)((nw + WMAX - 1) / WMAX);
        const unsigned max_seg = __reduce_max_sync(FULL, nseg);
        const bool staged = s_staged[warp][stage] != 0;
        bsrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.b[warp][stage][0]) - s_bw0[warp][stage] : a.bases;
        const unsigned long long* msrc = staged ? reinterpret_cast<const unsigned long long*>(&stg.m[warp][stage][0]) - s_mw0[warp][stage] : a.nmask;
        __syncwarp();

        for (unsigned seg = 0; seg < max_seg; ++seg) {
            const unsigned c = seg < nseg ? (unsigned)((nw - (unsigned long long)seg * WMAX) < WMAX ? (nw - (unsigned long long)seg * WMAX) : WMAX) : 0u;
            const unsigned long long s = R0 + (unsigned long long)seg * WMAX;

            uint32_t loc[SEGW];
            uint32_t nl[5];
#pragma unroll
            for (int k = 0; k < (int)SEGW; ++k) loc[k] = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) nl[k] = 0;
            if (c) {
                const unsigned long long q = s >> 5;
                const unsigned sh = 2u * (unsigned)(s & 31ull);
                unsigned long long W[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    unsigned long long idx = q + k;
                    if (idx >= a.base_words) idx = a.base_words - 1;
                    W[k] = bswap64(bsrc[idx]);
                }
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const unsigned long long v = sh ? ((W[k] << sh) | (W[k + 1] >> (64 - sh))) : W[k];
                    loc[2 * k] = (uint32_t)(v >> 32); loc[2 * k + 1] = (uint32_t)v;
                }
                if (HAS_NMASK) {
                    const unsigned long long qn2 = s >> 6;
                    const unsigned shn = (unsigned)(s & 63ull);
                    unsigned long long M[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        unsigned long long idx = qn2 + k;
                        if (idx >= a.nmask_words) idx = a.nmask_words - 1;
                        M[k] = bswap64(msrc[idx]);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const unsigned long long v = shn ? ((M[k] << shn) | (M[k + 1] >> (64 - shn))) : M[k];
                        if (2 * k < 5) nl[2 * k] = (uint32_t)(v >> 32);
                        if (2 * k + 1 < 5) nl[2 * k + 1] = (uint32_t)v;
                    }
                }
            }
            uint32_t v0, v1, v2;
            {
                if (HAS_NMASK && __any_sync(FULL, (nl[0] | nl[1] | nl[2] | nl[3] | nl[4]) != 0u)) {
                    unsigned cover = 1;
                    while (cover * 2 <= K) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], cover);
                        nl[4] |= nl[4] << cover;
                        cover *= 2;
                    }
                    const unsigned rest = K - cover;
                    if (rest) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) nl[k] |= fsl(nl[k], nl[k + 1], rest);
                        nl[4] |= nl[4] << rest;
                    }
                }
                const uint32_t c0 = c >= 32 ? 0xFFFFFFFFu : (c ? ~(0xFFFFFFFFu >> c) : 0u);
                const uint32_t c1 = c >= 64 ? 0xFFFFFFFFu : (c > 32 ? ~(0xFFFFFFFFu >> (c - 32)) : 0u);
                const uint32_t c2 = c >= 96 ? 0xFFFFFFFFu : (c > 64 ? ~(0xFFFFFFFFu >> (c - 64)) : 0u);
                v0 = ~nl[0] & c0; v1 = ~nl[1] & c1; v2 = ~nl[2] & c2;
            }
            my_valid += __popc(v0) + __popc(v1) + __popc(v2);
            if (__all_sync(FULL, (v0 | v1 | v2) == 0u)) continue;

            // the lane's copy of the segment (minimizer halves are fetched from it by position)
#pragma unroll
            for (int k = 0; k < (int)SEGW; ++k) sts32(seq_base + (uint32_t)k * MZ_ROW, loc[k]);
            sts32(seq_base + SEGW * MZ_ROW, 0u);

            // reverse complement of the segment: the complement of base x sits at base 154 - x of rcl[]
            uint32_t rcl[SEGW];
            {
                uint32_t t160[SEGW + 1];
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) t160[k] = rev2_32(~loc[SEGW - 1 - k]);
                t160[SEGW] = 0;
                constexpr unsigned bs = 2u * (160u - (WMAX + K - 1u));
                static_assert(bs < 32, "alignment shift must stay inside one word");
#pragma unroll
                for (int k = 0; k < (int)SEGW; ++k) rcl[k] = fsl(t160[k], t160[k + 1], bs);
            }

            // minimizer state carried from block to block: P = prefix minimum of position block 1 up to index 11
            uint32_t P = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < 12; ++i) P = min(P, mz_val(loc, rcl, 1, i));
            // the value of the run the lane is in, re-based to the coming block (positions are block-relative), and
            // whether its bit is set
            uint32_t held_wm = 0;
            bool have = false, held_pass = false;

            // Block b's runs are found by phase A in pass b; their bit-array words are loaded at the START of pass b + 1 and
            // looked at after that pass's phase A, which hides the DRAM latency.  Loads and their use sit in the same loop
            // body (nothing is in flight across the back edge), and no vote or other convergence point lies between them.
            bool pend = false, pneed0 = false, more = true;          // more: the segment has valid windows (checked above)
            uint32_t pW = 0, pchg = 0, pvm = 0, pnr = 0, pblk = 0;   // pW: positions of the minimizers of runs 0..3 (8 bits each)
            uint32_t pX0 = 0, pX1 = 0, pX2 = 0, pX3 = 0, pbits = 0, pbit2 = 0;  // word index and the two bits (8 bits each) of their lookups
#pragma unroll 1
            for (int blk = 0;; ++blk) {
                const uint32_t blk16 = (uint32_t)blk * 16u;
                // ---- load the bit-array words of the previous block's runs 0..3
                uint32_t F0 = 0, F1 = 0, F2 = 0, F3 = 0;
                if (pend) {
                    F0 = ldg_bitmap_if(pneed0, MB + pX0);
                    F1 = ldg_bitmap_if(pnr > 1u, MB + pX1);
                    F2 = ldg_bitmap_if(pnr > 2u, MB + pX2);
                    F3 = ldg_bitmap_if(pnr > 3u, MB + pX3);
                    my_fetch += (pneed0 ? 1u : 0u) + pnr - 1u;
                }
                // ---- phase A of this block: window minima and runs of equal minimizer (validity is ignored here: a window
                //      that is not valid costs at most a wasted lookup; it is masked out of the items)
                uint32_t chgraw = 0, wm_first = 0, vb = 0;
                if (more) {
                    vb = v0 & 0xFFFF0000u;
                    v0 = fsl(v0, v1, 16); v1 = fsl(v1, v2, 16); v2 <<= 16;
                    uint32_t Suf[16];
                    {
                        uint32_t smn = 0xFFFFFFFFu;
#pragma unroll
                        for (int i = 15; i >= 0; --i) { smn = min(smn, mz_val(loc, rcl, 0, i)); Suf[i] = smn; }
                    }
                    uint32_t A1 = 0, pw = 0;
                    uint32_t lst = wm_base + MZ_ROW;              // where the next run's value goes
#pragma unroll
                    for (int tt = 0; tt < 16; ++tt) {
                        uint32_t wm;
                        if (tt < 4) {                              // window tt: positions tt..15, then block 1 up to index tt + 12
                            P = min(P, mz_val(loc, rcl, 1, 12 + tt));
                            wm = min(Suf[tt], P);
                            if (tt == 3) { A1 = P; P = 0xFFFFFFFFu; }
                        } else {                                   // positions tt..15, all of block 1, block 2 up to index tt - 4
                            P = min(P, mz_val(loc, rcl, 2, tt - 4));
                            wm = min(min(Suf[tt], A1), P);
                        }
                        if (tt == 0) wm_first = wm;
                        else if (wm != pw) {                       // a new run starts at window tt
                            sts32(lst, wm);
                            lst += MZ_ROW;
                            chgraw |= 1u << tt;
                        }
                        pw = wm;
                    }
                    P -= 16u;                                      // block 2 of this block is block 1 of the next one
                }
                // ---- this block's runs: 0..3 are looked up (addresses now, loads at the start of the next pass, use after
                //      the next phase A); windows of later runs (rare) become items unfiltered
                bool nneed0 = false;
                uint32_t nW = 0, nchg = 0, nvm = 0, nnr = 1, nX0 = 0, nX1 = 0, nX2 = 0, nX3 = 0, nbits = 0, nbit2 = 0;
                if (more) {
                    uint32_t t3 = chgraw;
                    t3 &= t3 - 1u; t3 &= t3 - 1u; t3 &= t3 - 1u;      // changes beyond the third
                    nchg = chgraw ^ t3;
                    const uint32_t ovfm = t3 ? (~((t3 & (0u - t3)) - 1u) & 0xFFFFu) : 0u;   // every window from the fourth change on
                    const uint32_t vwin = __brev(vb) & 0xFFFFu;      // bit tt = validity of window tt
                    nnr = 1u + __popc(nchg);
                    nvm = vwin & ~ovfm;
                    if (__any_sync(FULL, t3 != 0u)) {
                        const uint32_t ia = lane | ((seg * WMAX + blk16) << 5);
                        uint32_t rest = t3;
                        unsigned rr = MZ_MAXRUN;
                        while (__any_sync(FULL, rest != 0u)) {
                            const uint32_t low = rest & (0u - rest), nxt = rest ^ low;
                            const uint32_t upto = nxt ? (nxt & (0u - nxt)) : 0x10000u;
                            const uint32_t mk = rest ? ((upto - low) & vwin) : 0u;
                            const uint32_t w = rest ? lds32(wm_base + rr * MZ_ROW) : 0u;
                            push(mk != 0u, ia, mk, (w & 63u) | 0x8000u);     // bit 31 of the item: level 1 not looked at yet
                            drain_if_full();
                            rest = nxt; ++rr;
                        }
                    }
                    const uint32_t W1 = lds32(wm_base + 1u * MZ_ROW), W2 = lds32(wm_base + 2u * MZ_ROW), W3 = lds32(wm_base + 3u * MZ_ROW);
                    nneed0 = !have || wm_first != held_wm;
                    have = true;
                    held_wm = (nnr == 1u ? wm_first : nnr == 2u ? W1 : nnr == 3u ? W2 : W3) - 16u;
                    nW = (wm_first & 63u) | ((W1 & 63u) << 8) | ((W2 & 63u) << 16) | ((W3 & 63u) << 24);
                    auto locate = [&](uint32_t w, unsigned sl) -> uint32_t {
                        uint32_t ha, hb;
                        halves_at(blk16 + (w & 63u), ha, hb);
                        const uint32_t zlo = mz_ident_lo(ha, hb) & fmask_lo, zhi = mz_ident_hi(ha, hb);
                        nbits |= (zlo & 31u) << (8u * sl);
                        nbit2 |= mz_bit2(zhi) << (8u * sl);
                        return (zlo >> 5) | ((zhi & fmask_hi) << 27);
                    };
                    nX0 = locate(wm_first, 0); nX1 = locate(W1, 1); nX2 = locate(W2, 2); nX3 = locate(W3, 3);
                }
                // ---- the previous block's words have had a phase A and the address work above to arrive
                if (pend) {
                    auto both = [&](uint32_t f, unsigned sh) -> bool {       // both bits of the identity set in its word
                        return ((f >> ((pbits >> sh) & 31u)) & (f >> ((pbit2 >> sh) & 31u)) & 1u) != 0u;
                    };
                    const bool b0 = pneed0 ? both(F0, 0) : held_pass;
                    const bool b1 = (pnr > 1u) & both(F1, 8);
                    const bool b2 = (pnr > 2u) & both(F2, 16);
                    const bool b3 = (pnr > 3u) & both(F3, 24);
                    held_pass = pnr == 1u ? b0 : pnr == 2u ? b1 : pnr == 3u ? b2 : b3;
                    // windows of run 0..3: [0, c1) [c1, c2) [c2, c3) [c3, 16); sentinels above bit 15 stand in for missing changes
                    uint32_t cc = pchg | 0x70000u;
                    const uint32_t c1 = cc & (0u - cc); cc ^= c1;
                    const uint32_t c2 = cc & (0u - cc); cc ^= c2;
                    const uint32_t c3 = cc & (0u - cc);
                    const uint32_t m0 = (c1 - 1u) & pvm, m1 = (c2 - c1) & pvm, m2 = (c3 - c2) & pvm, m3 = (0x10000u - c3) & pvm;
                    const bool q0 = b0 & (m0 != 0u), q1 = b1 & (m1 != 0u), q2 = b2 & (m2 != 0u), q3 = b3 & (m3 != 0u);
                    if (__any_sync(FULL, q0 | q1 | q2 | q3)) {
                        const uint32_t ia = lane | ((seg * WMAX + pblk) << 5);
                        push(q0, ia, m0, pW & 63u);
                        push(q1, ia, m1, (pW >> 8) & 63u);
                        push(q2, ia, m2, (pW >> 16) & 63u);
                        push(q3, ia, m3, pW >> 24);
                        drain_if_full();
                    }
                }
                if (!more) break;
                pneed0 = nneed0; pW = nW; pnr = nnr; pchg = nchg; pvm = nvm; pblk = blk16; pend = true;
                pX0 = nX0; pX1 = nX1; pX2 = nX2; pX3 = nX3; pbits = nbits; pbit2 = nbit2;
                more = blk + 1 < (int)(WMAX / 16) && !__all_sync(FULL, (v0 | v1 | v2) == 0u);   // anything valid after this block?
                // slide the register windows by one word
#pragma unroll
                for (int k = 0; k < (int)SEGW - 1; ++k) loc[k] = loc[k + 1];
#pragma unroll
                for (int k = (int)SEGW - 1; k > 0; --k) rcl[k] = rcl[k - 1];
            }
        }
        // items point into this tile's staged bases: finish them before the stage is refilled
        drain();
        const unsigned long long t_new = next_tile();
        if (lane == 0 && t_new < ntiles) issue(stage, t_new);
        t = t_ahead; t_ahead = t_new;
    }

    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(FULL, my_valid, o);
    my_fetch = __reduce_add_sync(FULL, my_fetch);
    if (lane == 0 && my_valid) atomicAdd(a.n_kmers, my_valid);
    if (lane == 0 && my_fetch) atomicAdd(a.n_kmers + 1, (unsigned long long)my_fetch);
}
constexpr size_t K1MZ_SMEM = sizeof(MzShared);

template <bool HAS_NMASK>
int launch_probe_mz(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    auto kern = k1_minimizer_probe<HAS_NMASK>;
    static bool done[64] = {};          // the attribute is per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1MZ_SMEM));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    CUDA_TRY(cudaMemsetAsync(a.tile_counter, 0, 8, st));
    kern<<<grid, RT, K1MZ_SMEM, st>>>(a, db);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
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
    if (db.layout == 2) {
        if (db.K != SK_K || !db.F) { mlg_set_error("minimizer-bitmap layout needs K=60 and its bitmap"); return MLG_ERR_STATE; }
        // persistent: the CTAs that fit pull 32-read tiles from a global counter
        const unsigned long long wtiles = (a.r_end - a.r_begin + 31) / 32;
        const unsigned long long need = (wtiles + WARPS - 1) / WARPS;
        int per_sm = MZ_MINCTAS;
        if (const char* e = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(e); if (x >= 1 && x <= 64) per_sm = x; }
        const unsigned long long res = (unsigned long long)ctx->sm_count * (unsigned)per_sm;
        grid = (unsigned)(need < res ? need : res);
        return a.nmask ? launch_probe_mz<true>(db, a, st, grid) : launch_probe_mz<false>(db, a, st, grid);
    }
    if (db.layout == 1) {
        if (db.K != SK_K || db.slots != 8) { mlg_set_error("super-k-mer layout needs K=60 and 8-slot buckets"); return MLG_ERR_STATE; }
        // persistent: the CTAs that fit (2 per SM) pull 32-read tiles from a global counter
        const unsigned long long wtiles = (a.r_end - a.r_begin + 31) / 32, res = (unsigned long long)ctx->sm_count * K1_MINCTAS;
        const unsigned long long need = (wtiles + WARPS - 1) / WARPS;
        grid = (unsigned)(need < res ? need : res);
        if (const char* e = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(e); if (x >= 1 && x <= 64) { unsigned long long w = (unsigned long long)ctx->sm_count * x; grid = (unsigned)(need < w ? need : w); } }
        return a.nmask ? launch_probe_sk<true>(db, a, st, grid) : launch_probe_sk<false>(db, a, st, grid);
    }
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
