// K1: decode + canonicalise + probe + count.   Replaces `kmc -k60 -ci2 -cs3` followed by
// `kmc_tools simple <db> <reads> intersect` (scripts/select_db.py:50-56): every N-free K-long window
// of every read is canonicalised (min of forward / reverse complement, A<C<G<T) and its occurrence is
// counted -- but only for k-mers of the database set D, which is all the intersection keeps.
//
// Data flow per CTA (persistent, one 256-word tile = 16384 bases at a time, double buffered):
//   TMA bulk copies (cp.async.bulk + mbarrier) stage the 2-bit packed bases (16 B / word), the N mask and
//   the read-start mask (8 B / word each) of the tile plus one halo word into shared memory;
//   each thread owns one 64-base word: it seeds the forward / reverse-complement 2K-bit registers from
//   the halo word, then rolls them base by base; window validity (no N inside, no read boundary inside)
//   is one 64-bit mask computed by shift-or smearing of the two bit masks;
//   each valid window: 64-bit hash of the canonical key -> bucket -> ONE 32-byte (or 16-byte) load of
//   31-bit fingerprints.  No fingerprint match and no overflow flag (> 99.9 % of windows) -> done.
//   Otherwise the exact path compares the full key against the bucket's run of D and bumps the
//   saturating 8-bit occurrence counter of that database k-mer (32-bit CAS).
// HBM traffic per window: one 32-byte sector of the fingerprint table.
#include "mlg_internal.h"

namespace {

constexpr unsigned TILE = MLG_TILE_WORDS;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ unsigned long long bswap64(unsigned long long v) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}
template <int SLOTS> struct BucketVec;
template <> struct BucketVec<8> { uint32_t w[8]; };
template <> struct BucketVec<4> { uint32_t w[4]; };

__device__ __forceinline__ void load_bucket(const uint32_t* T1, unsigned long long b, BucketVec<8>& v) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
                 : "l"(T1 + b * 8));
}
__device__ __forceinline__ void load_bucket(const uint32_t* T1, unsigned long long b, BucketVec<4>& v) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])
                 : "l"(T1 + b * 4));
}
// does this bucket need the exact path for fingerprint fp?
template <int SLOTS>
__device__ __forceinline__ bool bucket_candidate(const BucketVec<SLOTS>& v, uint32_t fp) {
    bool hit = ((v.w[0] & 0x7FFFFFFFu) == fp) | ((v.w[0] >> 31) != 0);
#pragma unroll
    for (int s = 1; s < SLOTS; ++s) hit |= (v.w[s] == fp);
    return hit;
}

// saturating (255) increment of byte counter i via 32-bit CAS
__device__ __forceinline__ void bump_counter(unsigned char* cnt8, uint32_t i) {
    uint32_t* wp = reinterpret_cast<uint32_t*>(cnt8) + (i >> 2);
    const uint32_t sh = (i & 3u) * 8u;
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    while (((old >> sh) & 0xFFu) < 0xFFu) {
        uint32_t assumed = old;
        old = atomicCAS(wp, assumed, assumed + (1u << sh));
        if (old == assumed) break;
    }
}
// exact path: compare the full key against the bucket's run of D, bump the counter on a match
__device__ __forceinline__ void probe_exact(const key128* __restrict__ D_key, const uint32_t* __restrict__ bstart,
                                            unsigned char* cnt8, unsigned long long bucket, unsigned long long khi,
                                            unsigned long long klo) {
    uint32_t s = bstart[bucket], e = bstart[bucket + 1];
    for (uint32_t i = s; i < e; ++i) {
        key128 d = D_key[i];
        if (d.hi == khi && d.lo == klo) { bump_counter(cnt8, i); return; }
    }
}

// OR of (v >> d) for d = 0 .. span-1 over a 128-bit value; only the low 64 bits of the result are used
__device__ __forceinline__ unsigned long long smear_low64(unsigned long long hi, unsigned long long lo, unsigned span) {
    if (span == 0) return 0ull;
    unsigned cover = 1;
    while (cover * 2 <= span) {
        unsigned long long nlo = lo | ((lo >> cover) | (hi << (64 - cover)));   // cover < 64 here
        unsigned long long nhi = hi | (hi >> cover);
        lo = nlo; hi = nhi; cover *= 2;
    }
    unsigned rest = span - cover;   // < cover <= 32
    if (rest) lo |= (lo >> rest) | (hi << (64 - rest));
    return lo;
}

constexpr unsigned WARPS = TILE / 32;
constexpr unsigned QCAP = 64;          // per-warp queue of exact-path candidates (drained at >= 32)

// The exact path is rare (true hits, fingerprint collisions, overflowed buckets) but made of dependent
// DRAM loads; taken inline it would serialise a whole warp on one lane.  Candidates are instead queued
// per warp in shared memory and drained 32 at a time, one candidate per lane.
struct WarpQueue {
    unsigned long long hi[WARPS][QCAP];
    unsigned long long lo[WARPS][QCAP];
    unsigned n[WARPS];
};

__device__ __forceinline__ void queue_drain(WarpQueue& q, unsigned warp, unsigned lane, const DbView& db, unsigned char* cnt8) {
    __syncwarp();
    const unsigned n = q.n[warp];
    for (unsigned i = lane; i < n; i += 32) {
        const unsigned long long khi = q.hi[warp][i], klo = q.lo[warp][i];
        key128 c; c.hi = khi; c.lo = klo;
        probe_exact(db.D_key, db.bstart, cnt8, hash_bucket(key_hash(c), db.bbits), khi, klo);
    }
    __syncwarp();
    if (lane == 0) q.n[warp] = 0;
    __syncwarp();
}
// all 32 lanes call this; `cand` lanes append their key
__device__ __forceinline__ void queue_push(WarpQueue& q, unsigned warp, unsigned lane, bool cand, unsigned long long khi,
                                           unsigned long long klo, const DbView& db, unsigned char* cnt8) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
    if (m == 0) return;                                   // warp-uniform
    const unsigned base = q.n[warp];
    if (cand) {
        const unsigned i = base + __popc(m & ((1u << lane) - 1u));
        q.hi[warp][i] = khi; q.lo[warp][i] = klo;
    }
    __syncwarp();
    const unsigned total = base + __popc(m);
    if (lane == 0) q.n[warp] = total;
    __syncwarp();
    if (total >= 32) queue_drain(q, warp, lane, db, cnt8);
}

template <int SLOTS, bool HAS_NMASK, int KT, bool USE_FILTER>
__global__ void __launch_bounds__(TILE, 2) k1_decode_canon_probe(ProbeArgs a, DbView db) {
    __shared__ __align__(16) uint4 sb[2][TILE + 2];
    __shared__ __align__(16) unsigned long long sn[2][TILE + 2];
    __shared__ __align__(16) unsigned long long ss[2][TILE + 2];
    __shared__ __align__(8) unsigned long long mbar[2];
    __shared__ unsigned long long s_total;
    __shared__ WarpQueue wq;

    const unsigned tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const unsigned K = KT ? (unsigned)KT : db.K;
    const unsigned long long ntiles = (a.w_end - a.w_begin + TILE - 1) / TILE;
    const unsigned fshift = 64u - db.fbits, bshift = 64u - db.bbits;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_total = 0;
    }
    if (lane == 0) wq.n[warp] = 0;
    __syncthreads();

    auto issue = [&](unsigned stage, unsigned long long t) {
        const unsigned long long t0 = a.w_begin + t * TILE;
        const unsigned nv = (unsigned)((a.w_end - t0) < TILE ? (a.w_end - t0) : TILE);
        uint32_t bytes_b, bytes_m;
        const uint4* src_b; const unsigned long long *src_n, *src_s;
        unsigned dst_b, dst_m;
        if (t0 > 0) {
            src_b = a.bases + (t0 - 1); dst_b = 0; bytes_b = (nv + 1) * 16u;
            unsigned words = nv + 2; words += words & 1u;
            src_n = a.nmask + (t0 - 2); src_s = a.smask + (t0 - 2); dst_m = 0; bytes_m = words * 8u;
        } else {
            src_b = a.bases; dst_b = 1; bytes_b = nv * 16u;
            unsigned words = nv + (nv & 1u);
            src_n = a.nmask; src_s = a.smask; dst_m = 2; bytes_m = words * 8u;
        }
        mbar_expect_tx(&mbar[stage], bytes_b + bytes_m * (HAS_NMASK ? 2u : 1u));
        bulk_g2s(&sb[stage][dst_b], src_b, bytes_b, &mbar[stage]);
        bulk_g2s(&ss[stage][dst_m], src_s, bytes_m, &mbar[stage]);
        if (HAS_NMASK) bulk_g2s(&sn[stage][dst_m], src_n, bytes_m, &mbar[stage]);
    };

    const key128 kmask = key_mask(2 * K);
    const unsigned top_shift = 2 * (K - 1);   // position of the first base
    unsigned long long my_valid = 0;

    unsigned it = 0;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const unsigned stage = it & 1u, parity = (it >> 1) & 1u;
        if (it == 0 && tid == 0) {
            issue(0, t);
            if (t + gridDim.x < ntiles) issue(1, t + gridDim.x);
        }
        mbar_wait(&mbar[stage], parity);

        const unsigned long long w = a.w_begin + t * TILE + tid;
        const bool active = w < a.w_end;
        uint4 cur = make_uint4(0, 0, 0, 0), prv = make_uint4(0, 0, 0, 0);
        unsigned long long nm_cur = ~0ull, nm_prv = 0, sm_cur = 0, sm_prv = 0;
        if (active) {
            cur = sb[stage][tid + 1];
            sm_cur = bswap64(ss[stage][tid + 2]);
            nm_cur = HAS_NMASK ? bswap64(sn[stage][tid + 2]) : 0ull;
            if (w > 0) {
                prv = sb[stage][tid];
                sm_prv = bswap64(ss[stage][tid + 1]);
                if (HAS_NMASK) nm_prv = bswap64(sn[stage][tid + 1]);
            } else {
                nm_prv = ~0ull;   // nothing before the stream: every window reaching back is invalid
            }
            if (w == a.nwords - 1) {
                unsigned r = (unsigned)(a.nbases & 63ull);
                if (r) nm_cur |= (1ull << (64 - r)) - 1ull;   // bases past the end of the stream
            }
        }
        __syncthreads();   // everyone has copied its words out of this stage
        if (tid == 0 && t + 2ull * gridDim.x < ntiles) issue(stage, t + 2ull * gridDim.x);

        // window-end validity: no N in the K bases ending here, no read start in the last K-1 of them.
        // (threads past the end of the range carry nm_cur = all ones: nothing valid, but they keep in step
        //  with their warp for the collectives below)
        const unsigned long long inval = smear_low64(nm_prv, nm_cur, K) | smear_low64(sm_prv, sm_cur, K - 1);
        const unsigned long long vmask = ~inval;
        my_valid += __popcll(vmask);

        // MSB-first 64-bit halves: bases 0..31 and 32..63 of the word
        const unsigned long long cur_hi = ((unsigned long long)__byte_perm(cur.x, 0, 0x0123) << 32) | __byte_perm(cur.y, 0, 0x0123);
        const unsigned long long cur_lo = ((unsigned long long)__byte_perm(cur.z, 0, 0x0123) << 32) | __byte_perm(cur.w, 0, 0x0123);
        key128 fwd;
        fwd.hi = ((unsigned long long)__byte_perm(prv.x, 0, 0x0123) << 32) | __byte_perm(prv.y, 0, 0x0123);
        fwd.lo = ((unsigned long long)__byte_perm(prv.z, 0, 0x0123) << 32) | __byte_perm(prv.w, 0, 0x0123);
        fwd = key_and(fwd, kmask);          // the K bases that precede this word
        key128 rcv = key_rc(fwd, K);

        constexpr int GROUP = 4;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const unsigned long long word = half ? cur_lo : cur_hi;
            const uint32_t vhalf = half ? (uint32_t)vmask : (uint32_t)(vmask >> 32);
            if (__all_sync(0xFFFFFFFFu, vhalf == 0u)) {
                // no lane of this warp has a valid window ending in this half: just advance the rolling registers
                if (half == 0) {
                    key128 nf;                       // (fwd << 64 | word) masked to 2K bits
                    nf.hi = fwd.lo; nf.lo = word;
                    fwd = key_and(nf, kmask);
                    rcv = key_rc(fwd, K);
                }
                continue;
            }
#pragma unroll 2
            for (int g0 = 0; g0 < 32; g0 += GROUP) {
                unsigned long long chi[GROUP], clo[GROUP], h[GROUP];
                bool ok[GROUP];
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
                    const int i = g0 + j;
                    const unsigned long long b = (word >> (62 - 2 * i)) & 3ull;
                    fwd.hi = (fwd.hi << 2) | (fwd.lo >> 62);
                    fwd.lo = (fwd.lo << 2) | b;
                    fwd = key_and(fwd, kmask);
                    rcv.lo = (rcv.lo >> 2) | (rcv.hi << 62);
                    rcv.hi = rcv.hi >> 2;
                    if (top_shift >= 64) rcv.hi |= (3ull - b) << (top_shift - 64);
                    else rcv.lo |= (3ull - b) << (top_shift & 63u);
                    ok[j] = (vhalf >> (31 - i)) & 1u;
                    const bool use_rc = key_lt(rcv, fwd);
                    chi[j] = use_rc ? rcv.hi : fwd.hi;
                    clo[j] = use_rc ? rcv.lo : fwd.lo;
                    key128 c; c.hi = chi[j]; c.lo = clo[j];
                    h[j] = key_hash(c);
                }
                if (USE_FILTER) {
                    // level 0: two bits of one 64-bit word of the L2-resident Bloom prefilter
                    unsigned long long fw[GROUP];
#pragma unroll
                    for (int j = 0; j < GROUP; ++j) {
                        fw[j] = 0ull;
                        if (ok[j]) fw[j] = __ldg(db.F + (h[j] >> fshift));
                    }
#pragma unroll
                    for (int j = 0; j < GROUP; ++j) {
                        const unsigned long long m = filter_mask(h[j]);
                        ok[j] = ok[j] && ((fw[j] & m) == m);
                    }
                }
                // level 1: one bucket of 31-bit fingerprints (a 32- or 16-byte sector of HBM)
                BucketVec<SLOTS> vec[GROUP];
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
#pragma unroll
                    for (int s = 0; s < SLOTS; ++s) vec[j].w[s] = 0u;
                    if (ok[j]) load_bucket(db.T1, db.bbits ? (h[j] >> bshift) : 0ull, vec[j]);
                }
#pragma unroll
                for (int j = 0; j < GROUP; ++j) {
                    const bool cand = ok[j] && bucket_candidate<SLOTS>(vec[j], hash_fp(h[j]));
                    queue_push(wq, warp, lane, cand, chi[j], clo[j], db, a.cnt8);
                }
            }
        }
    }
    queue_drain(wq, warp, lane, db, a.cnt8);

    // block-reduce the number of valid windows
    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xFFFFFFFFu, my_valid, o);
    if (lane == 0 && my_valid) atomicAdd(&s_total, my_valid);
    __syncthreads();
    if (tid == 0 && s_total) atomicAdd(a.n_kmers, s_total);
}

// ---------------------------------------------------------------- prep kernels
// store a logical MSB-first 64-bit mask in the stream's byte order (bit i -> byte i/8, bit 7-(i%8))
__device__ __forceinline__ unsigned long long to_stream_order(unsigned long long m) { return bswap64(m); }

__global__ void k_smask_fixed(unsigned long long* smask, unsigned long long nwords_alloc, unsigned long long nbases, uint32_t L) {
    unsigned long long w = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (w >= nwords_alloc) return;
    unsigned long long m = 0, base0 = w * 64ull;
    unsigned long long p = ((base0 + L - 1) / L) * L;
    for (; p < base0 + 64ull && p < nbases; p += L) m |= 1ull << (63 - (unsigned)(p - base0));
    smask[w] = to_stream_order(m);
}
__global__ void k_smask_offsets(unsigned long long* smask, const unsigned long long* off, unsigned long long nreads) {
    unsigned long long r = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    unsigned long long p = off[r];
    if (p >= off[nreads]) return;
    unsigned long long byte = p >> 3;
    uint32_t* wp = reinterpret_cast<uint32_t*>(smask) + (byte >> 2);
    atomicOr(wp, 1u << (8u * (unsigned)(byte & 3ull) + (7u - (unsigned)(p & 7ull))));
}
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

template <int SLOTS, bool HAS_NMASK, bool USE_FILTER>
int launch_probe_k(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    if (db.K == 60) k1_decode_canon_probe<SLOTS, HAS_NMASK, 60, USE_FILTER><<<grid, TILE, 0, st>>>(a, db);
    else k1_decode_canon_probe<SLOTS, HAS_NMASK, 0, USE_FILTER><<<grid, TILE, 0, st>>>(a, db);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
template <int SLOTS>
int launch_probe_s(const DbView& db, const ProbeArgs& a, cudaStream_t st, unsigned grid) {
    const bool nm = a.nmask != nullptr, fl = db.fbits != 0;
    if (nm) return fl ? launch_probe_k<SLOTS, true, true>(db, a, st, grid) : launch_probe_k<SLOTS, true, false>(db, a, st, grid);
    return fl ? launch_probe_k<SLOTS, false, true>(db, a, st, grid) : launch_probe_k<SLOTS, false, false>(db, a, st, grid);
}

}  // namespace

int launch_probe(const mlg_ctx* ctx, const DbView& db, const ProbeArgs& a, cudaStream_t st) {
    if (a.w_end <= a.w_begin) return MLG_OK;
    const unsigned long long ntiles = (a.w_end - a.w_begin + TILE - 1) / TILE;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        ctas_per_sm = 8;   // more CTAs than are resident: the hardware scheduler evens out the tail
        if (const char* s = getenv("MLG_PROBE_CTAS_PER_SM")) { int x = atoi(s); if (x >= 1 && x <= 64) ctas_per_sm = x; }
    }
    unsigned long long want = (unsigned long long)ctx->sm_count * ctas_per_sm;
    unsigned grid = (unsigned)(ntiles < want ? ntiles : want);
    return db.slots == 8 ? launch_probe_s<8>(db, a, st, grid) : launch_probe_s<4>(db, a, st, grid);
}

int launch_build_smask_fixed(unsigned long long* smask, unsigned long long nwords_alloc, unsigned long long nbases,
                             uint32_t read_len, cudaStream_t st) {
    if (!nwords_alloc) return MLG_OK;
    k_smask_fixed<<<(unsigned)((nwords_alloc + 255) / 256), 256, 0, st>>>(smask, nwords_alloc, nbases, read_len);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
}
int launch_build_smask_offsets(unsigned long long* smask, const unsigned long long* off, unsigned long long nreads,
                               cudaStream_t st) {
    if (!nreads) return MLG_OK;
    k_smask_offsets<<<(unsigned)((nreads + 255) / 256), 256, 0, st>>>(smask, off, nreads);
    CUDA_TRY(cudaGetLastError());
    return MLG_OK;
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
