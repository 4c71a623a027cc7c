// K1, layout 1 (MLG_LAYOUT=1): the round-1 super-k-mer kernel, kept as the A/B reference of probe_mz.cu.
#include "probe_common.cuh"

namespace {

#ifndef K1_MINCTAS
#define K1_MINCTAS 2                    // __launch_bounds__ minimum CTAs per SM (caps registers at 65536 / (256 * K1_MINCTAS))
#endif
constexpr unsigned QCAP = 64;           // per-warp queue of exact-path candidates (drained at >= 32)

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
constexpr unsigned SK_W = SK_K - MLG_MIN_M + 1;             // 45 minimizer positions per window
constexpr unsigned SK_MAXCH = 3;                             // new pairs fetched asynchronously per block of 16 windows
constexpr unsigned SK_NSLOT = SK_MAXCH + 1;                  // + the carried one
constexpr uint32_t SK_UNKNOWN = 0xFFFFFFFFu;
static_assert(SK_W == 45, "the block decomposition below is written for 45 positions");
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
            segment_load<HAS_NMASK>(loc, nl, c, s, bsrc, msrc, a.base_words, a.nmask_words);
            uint32_t v0, v1, v2;
            segment_valid<HAS_NMASK>(nl, c, K, v0, v1, v2);
            my_valid += __popc(v0) + __popc(v1) + __popc(v2);
            if (__all_sync(0xFFFFFFFFu, (v0 | v1 | v2) == 0u)) continue;

            // reverse complement of the segment: the complement of base x sits at base 154 - x of rcl[]
            uint32_t rcl[SEGW];
            segment_rc60(loc, rcl);

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
int launch_sk_t(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
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

}  // namespace

int launch_probe_sk(const DbView& db, const ProbeArgs& a, cudaStream_t st, int sm_count) {
    if (db.K != SK_K || db.slots != 8) { mlg_set_error("super-k-mer layout needs K=60 and 8-slot buckets"); return MLG_ERR_STATE; }
    // persistent: the CTAs that fit (2 per SM) pull 32-read tiles from a global counter
    const unsigned long long wtiles = (a.r_end - a.r_begin + 31) / 32;
    const unsigned long long need = (wtiles + WARPS - 1) / WARPS;
    int per_sm = K1_MINCTAS;
    if (const char* e = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(e); if (x >= 1 && x <= 64) per_sm = x; }
    const unsigned long long res = (unsigned long long)sm_count * (unsigned)per_sm;
    const unsigned grid = (unsigned)(need < res ? need : res);
    return a.nmask ? launch_sk_t<true>(db, a, st, grid) : launch_sk_t<false>(db, a, st, grid);
}
