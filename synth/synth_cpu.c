/* Synthetic workload generator, host build (OpenMP).  TEST / BENCH INFRASTRUCTURE.
 * See synth_core.h for the model.  Build: make -C synth  ->  synth/libmlg_synth_cpu.so
 */
#include <stdlib.h>
#include <string.h>
#include "synth_core.h"

#define SYN_API __attribute__((visibility("default")))

/* cum[i] = inclusive prefix sum of present-genome weights; length n_present */
SYN_API void syn_present_cum(const syn_params* p, uint64_t* cum) {
    uint64_t acc = 0;
    for (uint32_t i = 0; i < p->n_present; ++i) { acc += syn_present_weight(p, i); cum[i] = acc; }
}

/* keys: G*n pairs (hi, lo); empty slots are (~0, ~0) */
SYN_API void syn_gen_sketch_keys(const syn_params* p, uint64_t* keys) {
    uint64_t total = (uint64_t)p->G * p->n;
#pragma omp parallel for schedule(static, 4096)
    for (uint64_t s = 0; s < total; ++s) {
        uint32_t g = (uint32_t)(s / p->n), j = (uint32_t)(s % p->n);
        syn_sketch_key(p, g, j, &keys[2 * s], &keys[2 * s + 1]);
    }
}

/* ASCII reads r0 .. r0+nreads-1, L characters each, no separators */
SYN_API void syn_gen_reads_ascii(const syn_params* p, uint64_t r0, uint64_t nreads, char* out) {
    uint64_t* cum = (uint64_t*)malloc(sizeof(uint64_t) * p->n_present);
    syn_present_cum(p, cum);
    const uint32_t L = p->read_len;
#pragma omp parallel for schedule(static, 256)
    for (uint64_t i = 0; i < nreads; ++i) {
        uint64_t r = r0 + i;
        syn_read_src s = syn_read_source(p, cum, r);
        syn_gcache gc; syn_gcache_init(p, &gc, s.g);
        char* o = out + i * L;
        for (uint32_t t = 0; t < L; ++t) o[t] = "ACGTN"[syn_read_base(p, &s, &gc, r, t)];
    }
    free(cum);
}

/* 2-bit packed reads, back to back at base granularity (base i of the stream in byte i/4,
 * bits 7-2*(i%4)..6-2*(i%4)); nmask bit i in byte i/8, bit 7-(i%8).  N bases are packed as A.
 * bases: ceil(nreads*L/4) bytes rounded up to 16; nmask: ceil(nreads*L/8) rounded up to 16.
 * Both buffers must have been allocated with that rounding; pad bits are written as 0. */
SYN_API void syn_gen_reads_packed(const syn_params* p, uint64_t r0, uint64_t nreads,
                                  uint8_t* bases, uint8_t* nmask) {
    uint64_t* cum = (uint64_t*)malloc(sizeof(uint64_t) * p->n_present);
    syn_present_cum(p, cum);
    const uint32_t L = p->read_len;
    const uint64_t nb = nreads * L;
    const uint64_t nwords = (nb + 63) / 64;
#pragma omp parallel for schedule(static, 64)
    for (uint64_t w = 0; w < nwords; ++w) {
        uint8_t bb[16]; uint8_t mm[8];
        memset(bb, 0, 16); memset(mm, 0, 8);
        uint64_t cur_r = ~0ull; syn_read_src s; s.g = 0; s.start = 0; s.rev = 0;
        syn_gcache gc; syn_gcache_init(p, &gc, 0);
        for (uint32_t i = 0; i < 64; ++i) {
            uint64_t pos = w * 64 + i;
            if (pos >= nb) break;
            uint64_t r = r0 + pos / L; uint32_t t = (uint32_t)(pos % L);
            if (r != cur_r) { s = syn_read_source(p, cum, r); syn_gcache_init(p, &gc, s.g); cur_r = r; }
            uint32_t b = syn_read_base(p, &s, &gc, r, t);
            if (b == 4u) { mm[i >> 3] |= (uint8_t)(0x80u >> (i & 7)); b = 0; }
            bb[i >> 2] |= (uint8_t)(b << (6 - 2 * (i & 3)));
        }
        memcpy(bases + 16 * w, bb, 16);
        memcpy(nmask + 8 * w, mm, 8);
    }
    /* zero the alignment pad after the last word */
    uint64_t bbytes = ((nwords * 16 + 15) / 16) * 16, mbytes = ((nwords * 8 + 15) / 16) * 16;
    if (bbytes > nwords * 16) memset(bases + nwords * 16, 0, bbytes - nwords * 16);
    if (mbytes > nwords * 8) memset(nmask + nwords * 8, 0, mbytes - nwords * 8);
    free(cum);
}

SYN_API uint32_t syn_sizeof_params(void) { return (uint32_t)sizeof(syn_params); }
