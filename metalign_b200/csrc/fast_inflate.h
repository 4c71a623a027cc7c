// fast_inflate.h -- a DEFLATE (RFC 1951) / gzip (RFC 1952) decoder for the read ingest (ingest.cpp).
//
// zlib's inflate() delivers ~0.29 GB/s of FASTQ text on one core, which makes a gzip-compressed reads file -- the usual
// kind -- the slowest thing in a run by two orders of magnitude.  This decoder takes the same approach as the fast
// decoders around (64-bit bit buffer refilled 7 bytes at a time, one 11-bit table lookup per literal / length symbol with
// second-level tables for longer codes, 8-byte match copies) and exploits what the ingest can guarantee: the WHOLE
// compressed file is in memory (mapped), and the output buffer has slack behind its limit, so the inner loop needs no
// per-byte bounds checks.  Output is produced in caller-sized pieces: decoding stops at a symbol boundary once the
// output limit is passed and resumes in the next buffer, whose first bytes must be preceded by the last 32 KiB produced.
// Written from the RFCs; no code from zlib or libdeflate.  The ingest checks every member's CRC-32 and length, and
// MLGI_ZLIB=1 switches back to zlib (tests decode everything both ways).
#pragma once
#include <stdint.h>
#include <string.h>

namespace fastinf {

constexpr unsigned LL_BITS = 11, D_BITS = 8;          // primary table index widths
constexpr unsigned LL_SIZE = (1u << LL_BITS) + 2048, D_SIZE = (1u << D_BITS) + 1024;
constexpr unsigned SLACK = 320;                        // bytes a call may write past its output limit (one match + copy overshoot)
constexpr unsigned WINDOW = 32768;

// table entry: bits 0-3 code bits to consume, bits 4-7 extra bits (or second-level index width), bits 8-9 kind, bits 16-31 value
enum : uint32_t { K_LIT = 0u << 8, K_BASE = 1u << 8, K_EOB = 2u << 8, K_LINK = 3u << 8, K_MASK = 3u << 8 };
static inline uint32_t mk(uint32_t kind, uint32_t value, uint32_t bits, uint32_t extra) { return (value << 16) | kind | (extra << 4) | bits; }

static const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t CL_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Inflater {
    const uint8_t* in = nullptr;
    const uint8_t* in_end = nullptr;
    uint64_t bitbuf = 0;
    unsigned bitcnt = 0;
    enum State { MEMBER_HEADER, BLOCK_HEADER, STORED, HUFFMAN, TRAILER } state = MEMBER_HEADER;
    bool last_block = false;
    uint32_t stored_left = 0;
    uint32_t crc_expected = 0, isize_expected = 0;     // of the member that just ended (return value 1)
    const char* error = nullptr;
    uint32_t ll[LL_SIZE], dd[D_SIZE];

    void refill() {
        if (in_end - in >= 8) {
            uint64_t w;
            memcpy(&w, in, 8);
            bitbuf |= w << bitcnt;
            in += (63 - bitcnt) >> 3;
            bitcnt |= 56;
        } else {
            while (bitcnt <= 56 && in < in_end) { bitbuf |= (uint64_t)*in++ << bitcnt; bitcnt += 8; }
        }
    }
    uint32_t bits(unsigned n) { const uint32_t v = (uint32_t)(bitbuf & ((1ull << n) - 1)); bitbuf >>= n; bitcnt -= n; return v; }
    bool need(unsigned n) { if (bitcnt < n) refill(); return bitcnt >= n; }

    // canonical Huffman decode table from code lengths; kind_of(sym) supplies the entry for a symbol
    template <typename F>
    bool build(uint32_t* table, unsigned size, unsigned tbits, const uint8_t* lens, unsigned nsyms, F entry_of) {
        unsigned count[16] = {0};
        for (unsigned s = 0; s < nsyms; ++s) ++count[lens[s]];
        count[0] = 0;
        uint32_t next[16];
        uint32_t code = 0;
        long left = 1;
        for (unsigned l = 1; l <= 15; ++l) {
            left = left * 2 - (long)count[l];
            if (left < 0) return false;                       // over-subscribed
            code = (code + count[l - 1]) << 1;
            next[l] = code;
        }
        memset(table, 0, sizeof(uint32_t) * size);             // 0 = invalid (bits field 0): hit only by incomplete codes
        // second-level tables: width per primary prefix = longest code with that prefix, minus tbits
        uint8_t sub_bits[1u << LL_BITS];
        memset(sub_bits, 0, 1u << tbits);
        uint32_t rev_of[320];
        {
            uint32_t nx[16];
            memcpy(nx, next, sizeof(nx));
            for (unsigned s = 0; s < nsyms; ++s) {
                const unsigned l = lens[s];
                if (!l) continue;
                uint32_t c = nx[l]++, r = 0;
                for (unsigned i = 0; i < l; ++i) r |= ((c >> i) & 1u) << (l - 1 - i);
                rev_of[s] = r;
                if (l > tbits) { const unsigned p = r & ((1u << tbits) - 1u); if (l - tbits > sub_bits[p]) sub_bits[p] = (uint8_t)(l - tbits); }
            }
        }
        unsigned free_at = 1u << tbits;
        for (unsigned p = 0; p < (1u << tbits); ++p)
            if (sub_bits[p]) {
                if (free_at + (1u << sub_bits[p]) > size) return false;
                table[p] = mk(K_LINK, free_at, tbits, sub_bits[p]);
                free_at += 1u << sub_bits[p];
            }
        for (unsigned s = 0; s < nsyms; ++s) {
            const unsigned l = lens[s];
            if (!l) continue;
            const uint32_t r = rev_of[s];
            if (l <= tbits) {
                const uint32_t e = entry_of(s, l);
                for (uint32_t i = r; i < (1u << tbits); i += 1u << l) table[i] = e;
            } else {
                const unsigned p = r & ((1u << tbits) - 1u);
                const uint32_t base = table[p] >> 16, sb = sub_bits[p];
                const uint32_t e = entry_of(s, l - tbits);
                for (uint32_t i = r >> tbits; i < (1u << sb); i += 1u << (l - tbits)) table[base + i] = e;
            }
        }
        return true;
    }
    static uint32_t ll_entry(unsigned s, unsigned bits) {
        if (s < 256) return mk(K_LIT, s, bits, 0);
        if (s == 256) return mk(K_EOB, 0, bits, 0);
        if (s > 285) return 0;                                  // 286, 287: cannot occur in valid data
        return mk(K_BASE, LEN_BASE[s - 257], bits, LEN_EXTRA[s - 257]);
    }
    static uint32_t d_entry(unsigned s, unsigned bits) {
        if (s > 29) return 0;
        return mk(K_BASE, DIST_BASE[s], bits, DIST_EXTRA[s]);
    }
    bool fixed_tables() {
        uint8_t l[288], d[30];
        for (unsigned i = 0; i < 288; ++i) l[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
        memset(d, 5, sizeof(d));
        return build(ll, LL_SIZE, LL_BITS, l, 288, ll_entry) && build(dd, D_SIZE, D_BITS, d, 30, d_entry);
    }
    bool dynamic_tables() {
        if (!need(14)) return false;
        const unsigned hlit = bits(5) + 257, hdist = bits(5) + 1, hclen = bits(4) + 4;
        if (hlit > 286 || hdist > 30) return false;
        uint8_t cl[19] = {0};
        for (unsigned i = 0; i < hclen; ++i) { if (!need(3)) return false; cl[CL_ORDER[i]] = (uint8_t)bits(3); }
        uint32_t clt[1u << 7];
        // the code-length code has at most 7-bit codes: a 7-bit primary table holds it without second-level tables
        if (!build(clt, 1u << 7, 7, cl, 19, [](unsigned s, unsigned b) { return mk(K_LIT, s, b, 0); })) return false;
        uint8_t lens[288 + 32];
        unsigned n = 0;
        while (n < hlit + hdist) {
            if (!need(7 + 7)) { if (bitcnt == 0) return false; }
            const uint32_t e = clt[bitbuf & 127u];
            if ((e & 15u) == 0 || (e & 15u) > bitcnt) return false;
            bits(e & 15u);
            const unsigned sym = e >> 16;
            if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
            unsigned rep, val = 0;
            const unsigned xb = sym == 16 ? 2u : sym == 17 ? 3u : 7u;
            if (bitcnt < xb) return false;                       // truncated inside the repeat count
            if (sym == 16) { if (n == 0) return false; val = lens[n - 1]; rep = 3 + bits(2); }
            else if (sym == 17) rep = 3 + bits(3);
            else rep = 11 + bits(7);
            if (n + rep > hlit + hdist) return false;
            while (rep--) lens[n++] = (uint8_t)val;
        }
        if (lens[256] == 0) return false;                       // no end-of-block code
        return build(ll, LL_SIZE, LL_BITS, lens, hlit, ll_entry) && build(dd, D_SIZE, D_BITS, lens + hlit, hdist, d_entry);
    }
    bool member_header() {
        // byte-wise: the bit buffer is empty at a member boundary
        if (in_end - in < 18) return false;
        if (in[0] != 0x1f || in[1] != 0x8b || in[2] != 8) return false;
        const unsigned flg = in[3];
        const uint8_t* p = in + 10;
        if (flg & 4) { if (in_end - p < 2) return false; const unsigned xl = p[0] | (p[1] << 8); p += 2; if ((size_t)(in_end - p) < xl) return false; p += xl; }
        if (flg & 8) { while (p < in_end && *p) ++p; if (p >= in_end) return false; ++p; }
        if (flg & 16) { while (p < in_end && *p) ++p; if (p >= in_end) return false; ++p; }
        if (flg & 2) { if (in_end - p < 2) return false; p += 2; }
        in = p;
        return true;
    }

    // Decode into [out, limit + SLACK); `begin` = first byte that may be referenced by a match (>= 32 KiB of history
    // before out, or the start of the member's output).  Returns 0 when out has passed limit (call again with a fresh
    // buffer), 1 when a gzip member has ended (crc_expected / isize_expected are set), 2 at the end of the input, -1 on error.
    int run(uint8_t*& out, uint8_t* limit, const uint8_t* begin) {
        for (;;) {
            switch (state) {
            case MEMBER_HEADER:
                if (in >= in_end) return 2;
                // trailing zero padding after the last member is tolerated (some writers add it)
                if (*in == 0) { const uint8_t* p = in; while (p < in_end && *p == 0) ++p; if (p == in_end) { in = p; return 2; } }
                if (!member_header()) { error = "bad gzip member header"; return -1; }
                bitbuf = 0; bitcnt = 0;
                state = BLOCK_HEADER;
                break;
            case BLOCK_HEADER: {
                if (!need(3)) { error = "truncated deflate stream"; return -1; }
                last_block = bits(1) != 0;
                const unsigned type = bits(2);
                if (type == 0) {
                    bits(bitcnt & 7u);                                   // to the byte boundary
                    if (!need(32)) { error = "truncated stored block"; return -1; }
                    const unsigned len = bits(16), nlen = bits(16);
                    if ((len ^ nlen) != 0xFFFFu) { error = "bad stored block length"; return -1; }
                    stored_left = len;
                    state = STORED;
                } else if (type == 1) {
                    if (!fixed_tables()) { error = "internal: fixed tables"; return -1; }
                    state = HUFFMAN;
                } else if (type == 2) {
                    if (!dynamic_tables()) { error = "bad dynamic Huffman block header"; return -1; }
                    state = HUFFMAN;
                } else { error = "bad deflate block type"; return -1; }
                break;
            }
            case STORED: {
                // bytes still in the bit buffer first (whole bytes: the buffer was aligned), then straight from the input
                while (stored_left && bitcnt >= 8 && out < limit) { *out++ = (uint8_t)bits(8); --stored_left; }
                if (stored_left && bitcnt < 8) {
                    // give the unread whole bytes back: with bitcnt < 8 there are none, so the buffer is simply dropped
                    bitbuf = 0; bitcnt = 0;
                    size_t n = stored_left;
                    if ((size_t)(in_end - in) < n) { error = "truncated stored block"; return -1; }
                    if (out + n > limit) n = out < limit ? (size_t)(limit - out) : 0;
                    memcpy(out, in, n);
                    out += n; in += n; stored_left -= (uint32_t)n;
                }
                if (stored_left) return 0;
                state = last_block ? TRAILER : BLOCK_HEADER;
                break;
            }
            case HUFFMAN: {
                // one refill per symbol (measured: decoding several symbols per refill is slower here -- the refill is five
                // well-predicted instructions, the symbol-kind branches are what costs)
                for (;;) {
                    if (out >= limit) return 0;
                    refill();
                    uint32_t e = ll[bitbuf & ((1u << LL_BITS) - 1u)];
                    if ((e & K_MASK) == K_LINK) {
                        if (bitcnt < LL_BITS) { error = "truncated deflate stream"; return -1; }   // bitcnt is unsigned: never let it wrap
                        bitbuf >>= LL_BITS; bitcnt -= LL_BITS;
                        e = ll[(e >> 16) + (uint32_t)(bitbuf & ((1u << ((e >> 4) & 15u)) - 1u))];
                    }
                    const unsigned nb = e & 15u;
                    if (nb == 0 || nb > bitcnt) { error = bitcnt ? "invalid literal/length code" : "truncated deflate stream"; return -1; }
                    bitbuf >>= nb; bitcnt -= nb;
                    if ((e & K_MASK) == K_LIT) {
                        *out++ = (uint8_t)(e >> 16);
                        // a second literal from the same refill is the common case in text
                        const uint32_t e2 = ll[bitbuf & ((1u << LL_BITS) - 1u)];
                        if ((e2 & K_MASK) == K_LIT && (e2 & 15u) != 0 && bitcnt >= 24) {
                            bitbuf >>= (e2 & 15u); bitcnt -= (e2 & 15u);
                            *out++ = (uint8_t)(e2 >> 16);
                        }
                        continue;
                    }
                    if ((e & K_MASK) == K_EOB) break;
                    const unsigned xb = (e >> 4) & 15u;
                    if (xb > bitcnt) { error = "truncated deflate stream"; return -1; }
                    const unsigned len = (e >> 16) + (unsigned)(bitbuf & ((1u << xb) - 1u));
                    bitbuf >>= xb; bitcnt -= xb;
                    if (bitcnt < 32) refill();
                    uint32_t d = dd[bitbuf & ((1u << D_BITS) - 1u)];
                    if ((d & K_MASK) == K_LINK) {
                        if (bitcnt < D_BITS) { error = "truncated deflate stream"; return -1; }
                        bitbuf >>= D_BITS; bitcnt -= D_BITS;
                        d = dd[(d >> 16) + (uint32_t)(bitbuf & ((1u << ((d >> 4) & 15u)) - 1u))];
                    }
                    const unsigned db = d & 15u, dx = (d >> 4) & 15u;
                    if (db == 0 || db + dx > bitcnt) { error = bitcnt ? "invalid distance code" : "truncated deflate stream"; return -1; }
                    bitbuf >>= db; bitcnt -= db;
                    const unsigned dist = (d >> 16) + (unsigned)(bitbuf & ((1u << dx) - 1u));
                    bitbuf >>= dx; bitcnt -= dx;
                    if ((size_t)(out - begin) < dist) { error = "match distance beyond the start of the data"; return -1; }
                    const uint8_t* src = out - dist;
                    uint8_t* const end = out + len;
                    if (dist >= 8) {
                        do { memcpy(out, src, 8); out += 8; src += 8; } while (out < end);      // may overshoot into the slack
                    } else if (dist == 1) {
                        memset(out, *src, len);
                    } else {
                        do { *out++ = *src++; } while (out < end);
                    }
                    out = end;
                }
                state = last_block ? TRAILER : BLOCK_HEADER;
                break;
            }
            case TRAILER: {
                bits(bitcnt & 7u);
                if (!need(32)) { error = "truncated gzip trailer"; return -1; }
                crc_expected = bits(32);
                if (!need(32)) { error = "truncated gzip trailer"; return -1; }
                isize_expected = bits(32);
                // whole bytes left in the bit buffer belong to the next member: give them back
                in -= bitcnt >> 3;
                bitbuf = 0; bitcnt = 0;
                state = MEMBER_HEADER;
                return 1;
            }
            }
        }
    }
};

}  // namespace fastinf
