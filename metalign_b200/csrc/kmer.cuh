// 2K-bit k-mer arithmetic shared by every kernel (and unit-tested on the host, tests/test_kmer_math.py).
//
// A K-mer (K <= 63) is a 2K-bit integer with the FIRST base in the most significant position
// (A=0 C=1 G=2 T=3), so integer order == lexicographic order == KMC's canonical order
// (reference: `kmc -k60` at scripts/select_db.py:50).  Held as (hi, lo) uint64.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MLG_HD __host__ __device__ __forceinline__
#else
#define MLG_HD static inline
#endif

struct __attribute__((aligned(16))) key128 {
    unsigned long long hi, lo;
};

MLG_HD bool key_eq(const key128& a, const key128& b) { return a.hi == b.hi && a.lo == b.lo; }
MLG_HD bool key_lt(const key128& a, const key128& b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
MLG_HD bool key_is_empty(const key128& a) { return a.hi == ~0ull; }

// logical shifts of the 128-bit value, 0 <= s <= 127
MLG_HD key128 key_shr(const key128& a, unsigned s) {
    key128 r;
    if (s == 0) return a;
    if (s >= 64) { r.hi = 0; r.lo = a.hi >> (s - 64); }
    else { r.lo = (a.lo >> s) | (a.hi << (64 - s)); r.hi = a.hi >> s; }
    return r;
}
MLG_HD key128 key_shl(const key128& a, unsigned s) {
    key128 r;
    if (s == 0) return a;
    if (s >= 64) { r.lo = 0; r.hi = a.lo << (s - 64); }
    else { r.hi = (a.hi << s) | (a.lo >> (64 - s)); r.lo = a.lo << s; }
    return r;
}
// mask of the low `bits` bits, 0 <= bits <= 128
MLG_HD key128 key_mask(unsigned bits) {
    key128 m;
    if (bits >= 128) { m.hi = ~0ull; m.lo = ~0ull; }
    else if (bits >= 64) { m.lo = ~0ull; m.hi = (bits == 64) ? 0ull : ((1ull << (bits - 64)) - 1ull); }
    else { m.hi = 0; m.lo = (bits == 0) ? 0ull : ((1ull << bits) - 1ull); }
    return m;
}
MLG_HD key128 key_and(const key128& a, const key128& b) { key128 r; r.hi = a.hi & b.hi; r.lo = a.lo & b.lo; return r; }

// reverse the order of the 32 two-bit groups of a 64-bit word
MLG_HD unsigned long long rev2_64(unsigned long long x) {
#ifdef __CUDA_ARCH__
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
#else
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    return (x >> 32) | (x << 32);
#endif
}
// reverse complement of a k-base value (k <= 64)
MLG_HD key128 key_rc(const key128& a, unsigned k) {
    key128 full;                       // reverse all 64 groups of the complemented 128-bit value ...
    full.hi = rev2_64(~a.lo);
    full.lo = rev2_64(~a.hi);
    return key_shr(full, 128 - 2 * k); // ... then drop the (64-k) groups that came from the zero padding
}
MLG_HD key128 key_canon(const key128& a, unsigned K) {
    key128 r = key_rc(a, K);
    return key_lt(r, a) ? r : a;
}
// the k-base window starting `off` bases into a K-base value
MLG_HD key128 key_sub(const key128& x, unsigned K, unsigned off, unsigned k) {
    return key_and(key_shr(x, 2 * (K - off - k)), key_mask(2 * k));
}
// the leading k bases of a K-base value
MLG_HD key128 key_prefix(const key128& x, unsigned K, unsigned k) { return key_shr(x, 2 * (K - k)); }

// ---- hashing ------------------------------------------------------------------------------------
// The probe kernel never materialises the canonical k-mer on its fast path.  It hashes a STRAND-SYMMETRIC
// digest instead: the four 32-bit words of (forward + reverse complement), both top-aligned in 128 bits and
// added word by word.  Either strand gives the same digest, so the database hashes each canonical key the
// same way.  (The digest loses a little information -- e.g. two keys that differ by the same amount at two
// mirrored positions share it -- which can only cause a fingerprint false positive; the exact path compares
// full keys.)
MLG_HD unsigned long long hash_digest(unsigned s3, unsigned s2, unsigned s1, unsigned s0) {
    unsigned long long lo = ((unsigned long long)s1 << 32) | s0, hi = ((unsigned long long)s3 << 32) | s2;
    unsigned long long x = lo ^ (hi * 0x9E3779B97F4A7C15ull);
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}
// hash of a K-mer given in either orientation (bottom-aligned key, as stored in the database)
MLG_HD unsigned long long key_hash(const key128& x, unsigned K) {
    const key128 f = key_shl(x, 128 - 2 * K), r = key_shl(key_rc(x, K), 128 - 2 * K);
    return hash_digest((unsigned)(f.hi >> 32) + (unsigned)(r.hi >> 32), (unsigned)f.hi + (unsigned)r.hi,
                       (unsigned)(f.lo >> 32) + (unsigned)(r.lo >> 32), (unsigned)f.lo + (unsigned)r.lo);
}
// Tables have power-of-two sizes, indexed by the TOP bits of the hash: monotone in h, so entries sorted
// by hash are grouped by bucket.
MLG_HD unsigned long long hash_bucket(unsigned long long h, unsigned bbits) { return bbits ? (h >> (64 - bbits)) : 0ull; }
// L2-resident prefilter: a plain bit array of nfw 32-bit words, ONE bit per key
// (at the ~2-3 bits per key that fit in L2 for a 1.6e8-key database one probe bit is as good as two)
MLG_HD unsigned filter_word(unsigned long long h, unsigned nfw) {
#ifdef __CUDA_ARCH__
    return __umulhi((unsigned)(h >> 32), nfw);
#else
    return (unsigned)(((unsigned long long)(unsigned)(h >> 32) * nfw) >> 32);
#endif
}
MLG_HD unsigned filter_bit(unsigned long long h) { return ((unsigned)h >> 26) & 31u; }
// bits of the key inside its filter word: one bit, or two (fk == 2) when the filter has enough bits per key
MLG_HD unsigned filter_mask(unsigned long long h, unsigned fk) {
    unsigned m = 1u << (((unsigned)h >> 26) & 31u);
    if (fk == 2) m |= 1u << (((unsigned)h >> 21) & 31u);
    return m;
}
// 31-bit non-zero fingerprint (bit 31 of a bucket's first word is the overflow flag, 0 = empty slot)
MLG_HD unsigned hash_fp(unsigned long long h) {
    unsigned f = (unsigned)h & 0x7FFFFFFFu;
    return f ? f : 1u;
}

// ---- super-k-mer layout (K >= MLG_MIN_M) -----------------------------------------------------------
// Consecutive K-mers of a read share most of their bases.  In this layout the level-1 bucket of a K-mer is
// chosen by its MINIMIZER -- the smallest mixed value among its K-M+1 canonical M-mers (M = 16, so an M-mer is one
// 32-bit word) -- instead of by a hash of the whole K-mer.  A run of consecutive windows with the same
// minimizer (a "super-k-mer", ~23 windows on average at K = 60) probes ONE 32-byte bucket, fetched once; inside
// the bucket a K-mer is still recognised by a 31-bit fingerprint of its strand-symmetric digest.  The
// minimizer is strand-symmetric (canonical M-mers), so a K-mer and its reverse complement agree on the bucket.
#define MLG_MIN_M 16u
#define MLG_MIN_MULT 0x9E3779B1u      /* odd: x -> x*MULT + ADD is a bijection of 32-bit words (pseudo-random order,   */
#define MLG_MIN_ADD 0x7F4A7C15u       /* and poly-A (x = 0) is not the smallest value)                                */
#define MLG_BKT_MULT 0x85EBCA6Bu      /* odd: spreads the (small) window minima over the bucket index space           */

// reverse the order of the 16 two-bit groups of a 32-bit word
MLG_HD unsigned rev2_32h(unsigned x) {
#ifdef __CUDA_ARCH__
    x = __brev(x);
#else
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    x = (x >> 16) | (x << 16);
#endif
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}
// mixed value of the canonical form of a 16-mer given both strands
MLG_HD unsigned mmer_mix(unsigned f, unsigned r) { return (f < r ? f : r) * MLG_MIN_MULT + MLG_MIN_ADD; }
// minimizer value of a K-mer (bottom-aligned key), K >= 16
MLG_HD unsigned key_minimizer(const key128& x, unsigned K) {
    unsigned best = 0xFFFFFFFFu;
    for (unsigned p = 0; p + MLG_MIN_M <= K; ++p) {
        const unsigned f = (unsigned)key_shr(x, 2 * (K - MLG_MIN_M - p)).lo;
        const unsigned h = mmer_mix(f, rev2_32h(~f));
        best = h < best ? h : best;
    }
    return best;
}
MLG_HD unsigned minimizer_bucket_hash(unsigned wmin) { return wmin * MLG_BKT_MULT; }
// Fingerprint of a K-mer in this layout: a mix of the values of its FIRST and LAST canonical M-mers (the probe
// kernel has both at hand from the minimizer scan).  Swapping strands swaps the two, and the sum does not care.
// It tells a window from its shifted neighbours -- all a fingerprint has to do; the exact path compares full keys.
MLG_HD unsigned sk_fp(unsigned h_first, unsigned h_last) {
    const unsigned x = (h_first + h_last) * 0x7FEB352Du;      // the high bits of the product are the well-mixed ones
    return (x >> 1) | 1u;                                     // 31 bits, never 0 (0 = empty slot)
}
// Level-1 buckets come in PAIRS in this layout: the minimizer picks the pair (one 64-byte fetch per super-k-mer),
// bit 13 of the K-mer's own fingerprint picks the half it lives in (bit 13 = 8192 = the byte distance of the two
// halves in the probe kernel's shared-memory slots).  A half whose 8 slots are all taken counts as overflowed
// (there is no flag bit): its windows go to the exact compare.  K-mers that share a minimizer -- by descent
// or, with 16-base minimizers and >1e8 keys, by chance -- are thereby spread over 16 slots instead of 8.
MLG_HD unsigned sk_pair_index(unsigned wmin, unsigned bbits) { return minimizer_bucket_hash(wmin) >> (33u - bbits); }   // 2 <= bbits <= 31
MLG_HD unsigned sk_half(unsigned fp) { return (fp >> 13) & 1u; }
// 64-bit sort / bucket / fingerprint hash of a K-mer in the super-k-mer layout: the top bbits bits are the bucket
// (2 * pair + half; hash_bucket reads them), the low 31 bits are the fingerprint (hash_fp)
MLG_HD unsigned long long key_hash_sk(const key128& x, unsigned K, unsigned bbits) {
    unsigned best = 0xFFFFFFFFu, h_first = 0, h_last = 0;
    for (unsigned p = 0; p + MLG_MIN_M <= K; ++p) {
        const unsigned f = (unsigned)key_shr(x, 2 * (K - MLG_MIN_M - p)).lo;
        const unsigned h = mmer_mix(f, rev2_32h(~f));
        best = h < best ? h : best;
        if (p == 0) h_first = h;
        h_last = h;
    }
    const unsigned fp = sk_fp(h_first, h_last);
    const unsigned long long bucket = 2ull * sk_pair_index(best, bbits) + sk_half(fp);
    return (bucket << (64u - bbits)) | fp;
}

// ---- minimizer-bitmap layout (K >= MLG_MZ_M; db.layout == 2) -------------------------------------------
// With 16-base minimizers and >1e8 database K-mers the 2^32 minimizer values are saturated (window minima crowd the
// low 1/45 of the value range), so a minimizer alone says nothing about membership.  This layout uses 32-BASE
// minimizers, made of two 16-base halves so that everything stays in 32-bit words:
//   * for the 32-mer at position p of a K-mer let a = its first 16 bases and b = the REVERSE COMPLEMENT of its last 16
//     bases (both one word, first base most significant).  On the other strand the same 32-mer has (a, b) swapped.
//   * order: mz_order(a, b) = low 26 bits of (a + b) * MLG_MZ_ORD_MULT -- symmetric, so both strands agree.  The
//     minimizer of a K-mer is the 32-mer of smallest order, ties to the LEFTMOST position in reading direction.
//     The probe kernel computes (order << 6 | position) with two multiply-adds (the multiplier carries the << 6, the
//     position rides in the addend), so a plain unsigned min yields the minimum, the tie-break AND where it is.
//   * identity: mz_ident(a, b), a 64-bit mix of the unordered pair {a, b}, i.e. of the canonical 32-mer.  Level 1 is
//     a bit array indexed by the low bits of it, two bits per identity in one word: a run of windows sharing a
//     minimizer (a super-k-mer) with either bit clear contains no database K-mer, decided by ONE DRAM access and no
//     per-window compare.  2^64 identities do not saturate; false positives are the square of the bit density
//     (<= 1/16 by the sizing rule in db.cu, so ~0.2 %).
//   * a database K-mer x is filed under the identity of its leftmost minimum (what a read carrying x forward finds);
//     a read carrying rc(x) finds x's RIGHTMOST minimum.  The two differ only when the order ties between 32-mers of
//     different content (~4e-7 of the K-mers); those K-mers get a second bit and an entry in a small alias table.
#define MLG_MZ_M 32u
#define MLG_MZ_ORD_MULT 0x9E3779B1u
MLG_HD unsigned mz_order(unsigned a, unsigned b) { return ((a + b) * MLG_MZ_ORD_MULT) & 0x03FFFFFFu; }
// low word: bit index inside the level-1 array (plus the low bits of the high word for arrays above 2^32 bits);
// high word: bucket of the exact-compare index (its top bbits bits) and the second bit position (its top 5 bits).
// Both words are multiply-add hashes of the ordered pair (min, max) with one xor-shift each (10 instructions for the two;
// the round-2 version, with a multiply-xorshift finisher per word, took 17, four times per block of 16 windows).  Both
// words are used from their low bits up as well as from the top, hence the xor-shifts.  (A version hashing a + b and
// a ^ b instead of min / max lost a bit -- both products are even or odd together -- and with it half of the array:
// K1 went from 2.32 to 2.59 ms on four times the false positives.)
MLG_HD unsigned mz_ident_lo(unsigned a, unsigned b) {
    const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
    const unsigned t = lo * 0x85EBCA6Bu + hi * 0xC2B2AE35u;
    return t ^ (t >> 15);
}
MLG_HD unsigned mz_ident_hi(unsigned a, unsigned b) {
    const unsigned lo = a < b ? a : b, hi = a < b ? b : a;
    const unsigned u = lo * 0x27D4EB2Fu + hi * 0x165667B1u + 0x9E3779B9u;
    return u ^ (u >> 16);       // arrays above 2^32 bits take index bits from the LOW end of this word as well
}
MLG_HD unsigned long long mz_ident(unsigned a, unsigned b) { return ((unsigned long long)mz_ident_hi(a, b) << 32) | mz_ident_lo(a, b); }
// every identity sets TWO bits of its 32-bit word (blocked Bloom filter: one DRAM access, false positives ~ density^2):
// bit (ident & 31) and this one, taken from identity bits that do not take part in the word index
MLG_HD unsigned mz_bit2(unsigned ident_hi) { return ident_hi >> 27; }
MLG_HD unsigned long long mz_bit_index(unsigned long long ident, unsigned fbits) {       // 5 <= fbits <= 36
    return fbits >= 64 ? ident : (ident & ((1ull << fbits) - 1ull));
}
MLG_HD unsigned mz_bucket(unsigned long long ident, unsigned bbits) { return (unsigned)(ident >> 32) >> (32u - bbits); }   // 1 <= bbits <= 31
// identities of the leftmost and the rightmost minimum of a K-mer (bottom-aligned key), K >= 32
MLG_HD void key_mz(const key128& x, unsigned K, unsigned long long* zL, unsigned long long* zR) {
    unsigned best = 0xFFFFFFFFu, aL = 0, bL = 0, aR = 0, bR = 0;
    for (unsigned p = 0; p + MLG_MZ_M <= K; ++p) {
        const unsigned a = (unsigned)key_shr(x, 2 * (K - 16 - p)).lo;
        const unsigned c = (unsigned)key_shr(x, 2 * (K - 32 - p)).lo;
        const unsigned b = rev2_32h(~c);
        const unsigned o = mz_order(a, b);
        if (o < best) { best = o; aL = a; bL = b; aR = a; bR = b; }
        else if (o == best) { aR = a; bR = b; }
    }
    *zL = mz_ident(aL, bL);
    *zR = mz_ident(aR, bR);
}
// sort / bucket hash of a K-mer in this layout: the identity of its leftmost minimum (K-mers of one minimizer adjacent)
MLG_HD unsigned long long key_hash_mz(const key128& x, unsigned K) {
    unsigned long long zL, zR;
    key_mz(x, K, &zL, &zR);
    return zL;
}
