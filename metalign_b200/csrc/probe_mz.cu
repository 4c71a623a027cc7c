// MZ_RT: threads per CTA of this kernel alone (every warp is its own pipeline, so the CTA size only sets how the
// register file and shared memory are cut up); the other K1 kernels keep 256
#ifdef MZ_RT
#define MLG_RT MZ_RT
#endif
#include "probe_common.cuh"

namespace {

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
// reads per hand-out: one tile counter increment, one TMA issue and one mbarrier wait serve MZ_TILE_READS / 32 warp
// passes.  64 halves the share of the per-tile code (~300 instructions on one lane) but doubles the staging buffers:
// 2 x 108 KB of shared memory per SM leave 28 KB of L1 and the kernel ran 18 % SLOWER (2.56 vs 2.16 ms, r4c); 32 it is.
#ifndef MZ_TILE_READS
#define MZ_TILE_READS 32
#endif
constexpr unsigned MZ_TR = MZ_TILE_READS;
static_assert(MZ_TR % 32 == 0 && MZ_TR >= 32 && MZ_TR <= 128, "tile = whole warp passes");
constexpr unsigned MZ_STAGE_B = MZ_TR * 40 + 64;       // MZ_TR reads x 160 bases fit; longer reads are gathered from global
constexpr unsigned MZ_STAGE_M = MZ_TR * 20 + 64;
struct MzStage {
    __align__(16) unsigned char b[WARPS][2][MZ_STAGE_B];
    __align__(16) unsigned char m[WARPS][2][MZ_STAGE_M];
};
struct MzShared {
    MzStage stg;
    uint32_t wm[MZ_LIST][RT];                 // [r][thread]: value of the r-th run START of the current block (r >= 1)
    uint32_t seq[SEGW + 1][RT];               // [word][thread]: the lane's current segment (160 bases, top-aligned words)
    uint32_t qa[WARPS][MZ_QCAP];              // item: stream position (in bases) of the block's first window, low / high word
    uint32_t qc[WARPS][MZ_QCAP];
    uint32_t qb[WARPS][MZ_QCAP];              // item: windows of the block to compare exactly (bit tt = window tt) | base of the
                                              //       minimizer relative to the block's first window << 16
    uint32_t qoff[WARPS][32];                 // drain: windows before each item of the batch
    uint32_t qn[WARPS];                       // items waiting
};
constexpr uint32_t MZ_ROW = RT * 4;           // byte stride between rows of wm[] / seq[]

// The L2::64B qualifier matters: without a prefetch-size qualifier a miss on this part moves 128 bytes from HBM for
// every random 4-byte access (3.8 sectors per request, whatever the load's width or cache policy); with it, 64
// (profiles/r3a_ubench_gather.md).  MZ_L2_FETCH=0 builds the kernel without it (A/B measurements).
#ifndef MZ_L2_FETCH
#define MZ_L2_FETCH 64
#endif
#if MZ_L2_FETCH == 64
#define MZ_L2Q ".L2::64B"
#elif MZ_L2_FETCH == 128
#define MZ_L2Q ".L2::128B"
#else
#define MZ_L2Q ""
#endif
__device__ __forceinline__ uint32_t ldg_bitmap_if(bool p, const uint32_t* ptr) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %1, 0;\n"
        "mov.u32 %0, 0;\n"
        "@q ld.global.nc.L1::no_allocate" MZ_L2Q ".b32 %0, [%2];\n"
        "}\n" : "=r"(r) : "r"((uint32_t)p), "l"(ptr));
    return r;
}
// the exact path's random reads of database arrays, same fetch size
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t* ptr) {
    uint32_t r;
    asm volatile("ld.global.nc" MZ_L2Q ".b32 %0, [%1];" : "=r"(r) : "l"(ptr));
    return r;
}
__device__ __forceinline__ key128 ldg_key(const key128* ptr) {
    key128 k;
    asm volatile("ld.global.nc" MZ_L2Q ".v2.u64 {%0, %1}, [%2];" : "=l"(k.hi), "=l"(k.lo) : "l"(ptr));
    return k;
}
// The tile's bases and N bits as 32-bit shared loads from this warp's staging buffer (the common case: reads of up to 160
// bases; segment_load in probe_common.cuh is the general form -- 64-bit generic loads with clamped indices -- and cost 220
// of the ~850 instructions a read pays outside its blocks).  brel / mrel: the segment's first base relative to the first
// staged base / mask bit; sb / sk: shared addresses of the staged bytes (or of a zero block for a lane without windows).
template <bool HAS_NMASK>
__device__ __forceinline__ void segment_load_staged(uint32_t (&loc)[SEGW], uint32_t (&nl)[5], uint32_t brel, uint32_t mrel, uint32_t sb, uint32_t sk) {
    const uint32_t ab = sb + (brel >> 4) * 4u;
    const unsigned sh = 2u * (brel & 15u);
    uint32_t W[SEGW + 1];
#pragma unroll
    for (int k = 0; k <= (int)SEGW; ++k) W[k] = __byte_perm(lds32(ab + 4u * (uint32_t)k), 0, 0x0123);
#pragma unroll
    for (int k = 0; k < (int)SEGW; ++k) loc[k] = fsl(W[k], W[k + 1], sh);
#pragma unroll
    for (int k = 0; k < 5; ++k) nl[k] = 0;
    if (HAS_NMASK) {
        const uint32_t am = sk + (mrel >> 5) * 4u;
        const unsigned shn = mrel & 31u;
        uint32_t M[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) M[k] = __byte_perm(lds32(am + 4u * (uint32_t)k), 0, 0x0123);
#pragma unroll
        for (int k = 0; k < 5; ++k) nl[k] = fsl(M[k], M[k + 1], shn);
    }
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
__device__ __forceinline__ void mz_window_begin(MzWin& w, bool on, unsigned long long p, unsigned long long pm,
                                                const unsigned long long* bsrc, unsigned long long base_words, const DbView& db) {
    w.s = w.e = w.j0 = w.j1 = 0;
    w.cn.hi = w.cn.lo = 0; w.d0 = w.cn; w.d1 = w.cn;
    if (!on) return;
    // the window's 64 bases from p hold its minimizer too: it starts at most 28 bases in (pm - p <= K - 32)
    key128 F;
    mz_bases64(bsrc, base_words, p, F.hi, F.lo);
    const unsigned dsh = 2u * (unsigned)(pm - p);
    const unsigned long long mh = dsh ? ((F.hi << dsh) | (F.lo >> (64u - dsh))) : F.hi;
    const uint32_t ha = (uint32_t)(mh >> 32), hb = rev2_32(~(uint32_t)mh);
    const uint32_t zhi = mz_ident_hi(ha, hb), zlo = mz_ident_lo(ha, hb);
    const uint32_t bucket = zhi >> (32u - db.bbits);
    w.s = ldg_u32(db.bstart + bucket); w.e = ldg_u32(db.bstart + bucket + 1);
    // K-mers filed under a second identity (order ties) are rare: a 2^16-bit array says whether to look at all
    if (db.n_alias && ((db.alias_bloom[(zlo & 0xFFFFu) >> 5] >> (zlo & 31u)) & 1u)) {
        const unsigned long long z = ((unsigned long long)zhi << 32) | zlo;
        uint32_t lo = 0, hi = db.n_alias;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (db.alias_z[mid] < z) lo = mid + 1; else hi = mid; }
        w.j0 = w.j1 = lo;
        while (w.j1 < db.n_alias && db.alias_z[w.j1] == z) ++w.j1;
    }
    F = key_shr(F, 128 - 2 * SK_K);                                 // the 60-mer, bottom-aligned
    const key128 G = key_rc(F, SK_K);
    w.cn = key_lt(G, F) ? G : F;
    if (w.s < w.e) w.d0 = ldg_key(db.D_key + w.s);
    if (w.s + 1u < w.e) w.d1 = ldg_key(db.D_key + w.s + 1u);
}
// stage 2: compare with the database k-mers filed under the identity, count a match
__device__ __forceinline__ void mz_window_end(const MzWin& w, const DbView& db, const CountSink& cs) {
    if (w.s < w.e && w.d0.hi == w.cn.hi && w.d0.lo == w.cn.lo) { bump_counter(cs, w.s); return; }
    if (w.s + 1u < w.e && w.d1.hi == w.cn.hi && w.d1.lo == w.cn.lo) { bump_counter(cs, w.s + 1u); return; }
    for (uint32_t i = w.s + 2u; i < w.e; ++i) {
        const key128 d = ldg_key(db.D_key + i);
        if (d.hi == w.cn.hi && d.lo == w.cn.lo) { bump_counter(cs, i); return; }
    }
    for (uint32_t j = w.j0; j < w.j1; ++j) {
        const uint32_t i = db.alias_i[j];
        const key128 d = db.D_key[i];
        if (d.hi == w.cn.hi && d.lo == w.cn.lo) { bump_counter(cs, i); return; }
    }
}
// exact compare of every window of every waiting item of one warp, one WINDOW per lane (items have 1..16 windows)
// `a` and `db` are the kernel's __grid_constant__ parameters: the references point into parameter space, so nothing is
// copied to the stack for the call (before, every thread kept a 232-byte copy of both structs in local memory and the
// drain read its fields from there: 5.4 M local loads per 10 M reads, 18-27 % of them missing the L1).  Speed is the same
// within run-to-run noise (2.15 ms); what is left on the stack are two values parked around the block loop.
__device__ __noinline__ void mz_drain(uint32_t* qn, const uint32_t* qa, const uint32_t* qc, uint32_t* qb, uint32_t* qoff,
                                      const ProbeArgs& a, const DbView& db) {
    const unsigned long long* const bsrc = a.bases;
    const unsigned long long base_words = a.base_words;
    const CountSink cs{a.cnt8, a.present, a.n_present, a.touched, a.ci_min};
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    __syncwarp();
    const uint32_t n = *qn;
    for (uint32_t base = 0; base < n; base += 32u) {
        const uint32_t i = base + lane;
        uint32_t kb0 = i < n ? qb[i] : 0u;
        // Items of runs beyond the fourth of their block arrive untested (bit 31).  Their level-1 bits are looked at here,
        // once per ITEM and 32 items at a time, before the items are spread over the lanes window by window: 99.6 % of
        // them fail, and used to cost every one of their ~11 windows a lookup of its own.
        if (__any_sync(FULL, (kb0 >> 31) != 0u)) {
            if (kb0 >> 31) {
                const unsigned long long pb = ((unsigned long long)qc[i] << 32) | qa[i];
                unsigned long long mh, ml;
                mz_bases64(bsrc, base_words, pb + ((kb0 >> 16) & 63u), mh, ml);
                const uint32_t ha = (uint32_t)(mh >> 32), hb = rev2_32(~(uint32_t)mh);
                const uint32_t zhi = mz_ident_hi(ha, hb), zlo = mz_ident_lo(ha, hb);
                const unsigned long long idx = mz_bit_index(((unsigned long long)zhi << 32) | zlo, db.fbits);
                const uint32_t f = ldg_u32(db.F + (idx >> 5));
                kb0 = ((f >> (zlo & 31u)) & (f >> mz_bit2(zhi)) & 1u) ? (kb0 & 0x7FFFFFFFu) : 0u;
                qb[i] = kb0;
            }
            __syncwarp();
        }
        const uint32_t cnt = (uint32_t)__popc(kb0 & 0xFFFFu);
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
            const uint32_t kb = qb[base + j];
            const uint32_t tt = __fns(kb & 0xFFFFu, 0u, (int)(w - qoff[j]) + 1);
            const unsigned long long pb = ((unsigned long long)qc[base + j] << 32) | qa[base + j];
            MzWin win;
            mz_window_begin(win, true, pb + tt, pb + ((kb >> 16) & 63u), bsrc, base_words, db);
            mz_window_end(win, db, cs);
        }
        __syncwarp();
    }
    __syncwarp();                 // every lane has read *qn, also when the list was empty and the loop did not run
    if (lane == 0) *qn = 0;
    __syncwarp();
}

template <bool HAS_NMASK>
__global__ void __launch_bounds__(RT, MZ_MINCTAS) k1_minimizer_probe(const __grid_constant__ ProbeArgs a, const __grid_constant__ DbView db) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MzShared& sm = *reinterpret_cast<MzShared*>(smem_raw);
    MzStage& stg = sm.stg;
    __shared__ __align__(8) unsigned long long mbar[WARPS][2];
    __shared__ unsigned long long s_bw0[WARPS][2], s_mw0[WARPS][2];
    __shared__ unsigned s_staged[WARPS][2];
    __shared__ __align__(16) uint32_t s_zero[SEGW + 2];      // what a lane without windows loads instead of staged bytes

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    constexpr unsigned K = SK_K;
    if (tid < SEGW + 2) s_zero[tid] = 0u;
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const unsigned long long nreads = a.r_end - a.r_begin;
    const unsigned long long ntiles = (nreads + MZ_TR - 1) / MZ_TR;       // a tile = MZ_TR reads: MZ_TR / 32 warp passes
    // the hand-out counter: 32 bits are plenty (a tile is >= 32 reads), and the word above it stays zero
    unsigned int* const tile_ctr = reinterpret_cast<unsigned int*>(a.tile_counter);
    auto next_tile = [&]() -> unsigned long long {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(tile_ctr, 1u);
        return __shfl_sync(FULL, t, 0);
    };
    // The packed reads go through the L2 with NORMAL priority: the drain reads the bases of its items back from global memory
    // a tile or two later, and with evict_first (the round-1 choice, -DMZ_STREAM_EVICT_FIRST) the level-1 array's random lines
    // had pushed them out by then: 2.139 ms against 2.112 (evict_last: 2.112).
#ifdef MZ_STREAM_EVICT_FIRST
    const unsigned long long pol_stream = policy_evict_first();
#elif defined(MZ_STREAM_EVICT_LAST)
    const unsigned long long pol_stream = policy_evict_last();
#else
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_stream));
#endif
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
        const unsigned long long r0 = a.r_begin + t * (unsigned long long)MZ_TR;
        const unsigned long long r1 = (r0 + MZ_TR < a.r_end) ? r0 + MZ_TR : a.r_end;
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
        const bool fits = bw1 > bw0 && bytes_b <= MZ_STAGE_B && bytes_m <= MZ_STAGE_M;
        s_bw0[warp][stage] = bw0; s_mw0[warp][stage] = mw0; s_staged[warp][stage] = fits ? 1u : 0u;
        if (fits) {
            mbar_expect_tx(&mbar[warp][stage], (uint32_t)(bytes_b + bytes_m));
            bulk_g2s(&stg.b[warp][stage][0], a.bases + bw0, (uint32_t)bytes_b, &mbar[warp][stage], pol_stream);
            if (HAS_NMASK && bytes_m) bulk_g2s(&stg.m[warp][stage][0], a.nmask + mw0, (uint32_t)bytes_m, &mbar[warp][stage], pol_stream);
        } else {
            mbar_arrive(&mbar[warp][stage]);
        }
    };

    unsigned my_valid = 0;                    // per lane: a launch would need > 3e14 windows to overflow it
    unsigned my_fetch = 0;
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
        // items carry stream positions and the drain reads the bases from global memory (L2: they were just streamed in), so
        // items survive the tile they come from and a drain always has >= MZ_QDRAIN items' windows to spread over the lanes
        mz_drain(&sm.qn[warp], &sm.qa[warp][0], &sm.qc[warp][0], &sm.qb[warp][0], &sm.qoff[warp][0], a, db);
    };
    // lanes with p append one item: windows `ik` of the block whose first window is at stream base `pb`, minimizer at base `rel` of the block
    auto push = [&](bool p, unsigned long long pb, uint32_t ik, uint32_t rel) {
        if (p) {
            const uint32_t i = atomicAdd(&sm.qn[warp], 1u);
            sm.qa[warp][i] = (uint32_t)pb; sm.qc[warp][i] = (uint32_t)(pb >> 32); sm.qb[warp][i] = ik | (rel << 16);
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

        const bool staged = s_staged[warp][stage] != 0;
        const unsigned long long bw0_32 = s_bw0[warp][stage] * 32ull, mw0_64 = s_mw0[warp][stage] * 64ull;
        const uint32_t sb_stage = smem_u32(&stg.b[warp][stage][0]), sk_stage = smem_u32(&stg.m[warp][stage][0]), s_zero_a = smem_u32(&s_zero[0]);
#pragma unroll 1
        for (unsigned half = 0; half < MZ_TR / 32u; ++half) {
        const unsigned long long r = a.r_begin + t * (unsigned long long)MZ_TR + half * 32u + lane;
        if (half && r - lane >= a.r_end) break;
        const bool active = r < a.r_end;
        unsigned long long R0 = 0, R1 = 0;
        if (active) {
            R0 = a.off ? a.off[r] : r * (unsigned long long)a.read_len;
            R1 = a.off ? a.off[r + 1] : R0 + a.read_len;
        }
        const unsigned long long len = R1 - R0;
        const unsigned long long nw = len >= K ? len - K + 1 : 0ull;
        const unsigned nseg = (unsigned)((nw + WMAX - 1) / WMAX);
        const unsigned max_seg = __reduce_max_sync(FULL, nseg);
        // staged (reads of up to 160 bases): 32-bit positions relative to the staged bytes; otherwise the stream itself
        const uint32_t brel0 = (uint32_t)(R0 - bw0_32), mrel0 = (uint32_t)(R0 - mw0_64);
        __syncwarp();

        for (unsigned seg = 0; seg < max_seg; ++seg) {
            const unsigned c = seg < nseg ? (unsigned)((nw - (unsigned long long)seg * WMAX) < WMAX ? (nw - (unsigned long long)seg * WMAX) : WMAX) : 0u;
            const unsigned long long s = bw0_32 + brel0 + (unsigned long long)seg * WMAX;      // = R0 + seg * WMAX

            uint32_t loc[SEGW];
            uint32_t nl[5];
            if (staged) segment_load_staged<HAS_NMASK>(loc, nl, c ? brel0 + seg * WMAX : 0u, c ? mrel0 + seg * WMAX : 0u, c ? sb_stage : s_zero_a, c ? sk_stage : s_zero_a);
            else segment_load<HAS_NMASK>(loc, nl, c, s, a.bases, a.nmask, a.base_words, a.nmask_words);
            uint32_t v0, v1, v2;
            segment_valid<HAS_NMASK>(nl, c, K, v0, v1, v2);
            my_valid += __popc(v0) + __popc(v1) + __popc(v2);
            if (__all_sync(FULL, (v0 | v1 | v2) == 0u)) continue;

            // the lane's copy of the segment (minimizer halves are fetched from it by position)
#pragma unroll
            for (int k = 0; k < (int)SEGW; ++k) sts32(seq_base + (uint32_t)k * MZ_ROW, loc[k]);
            sts32(seq_base + SEGW * MZ_ROW, 0u);

            // reverse complement of the segment: the complement of base x sits at base 154 - x of rcl[]
            uint32_t rcl[SEGW];
            segment_rc60(loc, rcl);

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
                        const unsigned long long ia = s_bw0[warp][stage] * 32ull + brel0 + (unsigned long long)seg * WMAX + blk16;
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
                    // Few runs pass (0.4 % by chance, plus the true hits): in three blocks of four no lane has one, and the window
                    // masks below are not needed at all.
                    if (__any_sync(FULL, b0 | b1 | b2 | b3)) {
                        // windows of run 0..3: [0, c1) [c1, c2) [c2, c3) [c3, 16); sentinels above bit 15 stand in for missing changes
                        uint32_t cc = pchg | 0x70000u;
                        const uint32_t c1 = cc & (0u - cc); cc ^= c1;
                        const uint32_t c2 = cc & (0u - cc); cc ^= c2;
                        const uint32_t c3 = cc & (0u - cc);
                        const uint32_t m0 = (c1 - 1u) & pvm, m1 = (c2 - c1) & pvm, m2 = (c3 - c2) & pvm, m3 = (0x10000u - c3) & pvm;
                        const bool q0 = b0 & (m0 != 0u), q1 = b1 & (m1 != 0u), q2 = b2 & (m2 != 0u), q3 = b3 & (m3 != 0u);
                        const unsigned long long ia = s_bw0[warp][stage] * 32ull + brel0 + (unsigned long long)seg * WMAX + pblk;
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
        }
        __syncwarp();             // every lane is done with this stage's staged bases before it is refilled
        const unsigned long long t_new = next_tile();      // (asking for it at the start of the tile instead changes nothing: 2.164 ms either way, r4h)
        if (lane == 0 && t_new < ntiles) issue(stage, t_new);
        t = t_ahead; t_ahead = t_new;
    }

    drain();                      // the items still waiting
    unsigned long long warp_valid = my_valid;
    for (int o = 16; o > 0; o >>= 1) warp_valid += __shfl_down_sync(FULL, warp_valid, o);
    my_fetch = __reduce_add_sync(FULL, my_fetch);
    if (lane == 0 && warp_valid) atomicAdd(a.n_kmers, warp_valid);
    if (lane == 0 && my_fetch) atomicAdd(a.n_kmers + 1, (unsigned long long)my_fetch);
}

constexpr size_t K1MZ_SMEM = sizeof(MzShared);

template <bool HAS_NMASK>
int launch_mz_t(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    auto kern = k1_minimizer_probe<HAS_NMASK>;
    static bool done[64] = {};          // the attribute is per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K1MZ_SMEM));
        // MLG_PROBE_CARVEOUT=<percent of the SM's 228 KB>: shared-memory carve-out (experiments: what is left is the L1)
        if (const char* e = getenv("MLG_PROBE_CARVEOUT")) { int x = atoi(e); if (x >= 0 && x <= 100) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, x)); }
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    CUDA_TRY(cudaMemsetAsync(a.tile_counter, 0, 8, st));
    kern<<<grid, RT, K1MZ_SMEM, st>>>(a, db);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}

}  // namespace

int launch_probe_mz(const DbView& db, const ProbeArgs& a, cudaStream_t st, int sm_count) {
    if (db.K != SK_K || !db.F) { mlg_set_error("minimizer-bitmap layout needs K=60 and its bit array"); return MLG_ERR_STATE; }
    // persistent: the CTAs that fit pull 32-read tiles from a global counter
    const unsigned long long wtiles = (a.r_end - a.r_begin + MZ_TR - 1) / MZ_TR;
    const unsigned long long need = (wtiles + WARPS - 1) / WARPS;
    int per_sm = MZ_MINCTAS;
    if (const char* e = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(e); if (x >= 1 && x <= 64) per_sm = x; }
    const unsigned long long res = (unsigned long long)sm_count * (unsigned)per_sm;
    const unsigned grid = (unsigned)(need < res ? need : res);
    return a.nmask ? launch_mz_t<true>(db, a, st, grid) : launch_mz_t<false>(db, a, st, grid);
}
