// Device helpers shared by the three K1 probe kernels (probe.cu: whole-k-mer hash layout and the dispatcher;
// probe_sk.cu: fingerprint pairs by 16-base minimizer; probe_mz.cu: identity bit array by 32-base minimizer).
#pragma once
#include "mlg_internal.h"

namespace {

#ifndef MLG_RT
#define MLG_RT 256
#endif
constexpr unsigned RT = MLG_RT;         // reads per tile == threads per CTA
constexpr unsigned WARPS = RT / 32;
constexpr unsigned WMAX = 96;           // window starts per segment
constexpr unsigned SEGW = 10;           // 32-bit words of bases per segment (160 bases >= WMAX + 63 - 1)

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
// K = 60 kernels: every warp stages its own 32-read tiles
constexpr unsigned SK_K = 60;
constexpr unsigned WSTAGE_B = 1280 + 64;                     // per-warp staging: 32 reads x 160 bases fit; longer reads are gathered from global
constexpr unsigned WSTAGE_M = 640 + 64;

struct SkStage {
    __align__(16) unsigned char b[WARPS][2][WSTAGE_B];
    __align__(16) unsigned char m[WARPS][2][WSTAGE_M];
};
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

// ---- one segment of a read: WMAX window starts, 160 bases, in registers (shared by the three K1 kernels) ----
// gather the 160 bases (and their N bits) that start at stream base s, top-aligned; c = window starts of the segment (0: none)
template <bool HAS_NMASK>
__device__ __forceinline__ void segment_load(uint32_t (&loc)[SEGW], uint32_t (&nl)[5], unsigned c, unsigned long long s,
                                             const unsigned long long* bsrc, const unsigned long long* msrc,
                                             unsigned long long base_words, unsigned long long nmask_words) {
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
            if (idx >= base_words) idx = base_words - 1;
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
                if (idx >= nmask_words) idx = nmask_words - 1;
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
}
// validity of the 96 window starts (MSB first): i < c and no N in bases [i, i+K); nl[] is smeared in place
template <bool HAS_NMASK>
__device__ __forceinline__ void segment_valid(uint32_t (&nl)[5], unsigned c, unsigned K, uint32_t& v0, uint32_t& v1, uint32_t& v2) {
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
// K = 60: reverse complement of the segment; the complement of base x sits at base 154 - x of rcl[]
__device__ __forceinline__ void segment_rc60(const uint32_t (&loc)[SEGW], uint32_t (&rcl)[SEGW]) {
    uint32_t t160[SEGW + 1];
#pragma unroll
    for (int k = 0; k < (int)SEGW; ++k) t160[k] = rev2_32(~loc[SEGW - 1 - k]);
    t160[SEGW] = 0;
    constexpr unsigned bs = 2u * (160u - (WMAX + SK_K - 1u));       // 10 bits dropped at the front
    static_assert(bs < 32, "alignment shift must stay inside one word");
#pragma unroll
    for (int k = 0; k < (int)SEGW; ++k) rcl[k] = fsl(t160[k], t160[k + 1], bs);
}

}  // namespace

// launchers of the two K = 60 kernels (probe_sk.cu, probe_mz.cu); sm_count sizes the persistent grid
int launch_probe_sk(const DbView& db, const ProbeArgs& a, cudaStream_t st, int sm_count);
int launch_probe_mz(const DbView& db, const ProbeArgs& a, cudaStream_t st, int sm_count);
