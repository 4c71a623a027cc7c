// Synthetic workload generator, device build (sm_100a).  TEST / BENCH INFRASTRUCTURE -- see synth_core.h.
// Same functions as synth_cpu.c, writing into device memory.  Build: make -C synth
#include <cuda_runtime.h>
#include <stdio.h>
#include "synth_core.h"

#define SYN_API extern "C" __attribute__((visibility("default")))

__global__ void k_sketch_keys(syn_params p, uint64_t* keys) {
    uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)p.G * p.n;
    if (s >= total) return;
    uint64_t hi, lo;
    syn_sketch_key(&p, (uint32_t)(s / p.n), (uint32_t)(s % p.n), &hi, &lo);
    keys[2 * s] = hi; keys[2 * s + 1] = lo;
}

// one thread per 64-base word of the back-to-back packed stream
__global__ void k_reads_packed(syn_params p, const uint64_t* cum, uint64_t r0, uint64_t nreads, uint64_t nwords_alloc_b,
                               uint64_t nwords_alloc_m, uint8_t* bases, uint8_t* nmask) {
    uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint32_t L = p.read_len;
    const uint64_t nb = nreads * L;
    uint32_t bb[4] = {0, 0, 0, 0};   // 16 bytes, byte k = bits 8*(k%4) of bb[k/4]
    uint32_t mm[2] = {0, 0};
    if (w * 64 < nb) {
        uint64_t cur_r = ~0ull; syn_read_src s; s.g = 0; s.start = 0; s.rev = 0;
        syn_gcache gc; syn_gcache_init(&p, &gc, 0);
        for (uint32_t i = 0; i < 64; ++i) {
            uint64_t pos = w * 64 + i;
            if (pos >= nb) break;
            uint64_t r = r0 + pos / L; uint32_t t = (uint32_t)(pos % L);
            if (r != cur_r) { s = syn_read_source(&p, cum, r); syn_gcache_init(&p, &gc, s.g); cur_r = r; }
            uint32_t b = syn_read_base(&p, &s, &gc, r, t);
            if (b == 4u) { uint32_t byte = i >> 3; mm[byte >> 2] |= (0x80u >> (i & 7)) << (8 * (byte & 3)); b = 0; }
            uint32_t byte = i >> 2;
            bb[byte >> 2] |= (b << (6 - 2 * (i & 3))) << (8 * (byte & 3));
        }
    }
    if (w < nwords_alloc_b) reinterpret_cast<uint4*>(bases)[w] = make_uint4(bb[0], bb[1], bb[2], bb[3]);
    if (w < nwords_alloc_m) reinterpret_cast<uint2*>(nmask)[w] = make_uint2(mm[0], mm[1]);
}

static int fail(const char* what, cudaError_t e) { fprintf(stderr, "synth_cuda: %s: %s\n", what, cudaGetErrorString(e)); return -1; }

SYN_API int syn_cuda_gen_sketch_keys(const syn_params* p, void* d_keys, void* stream) {
    uint64_t total = (uint64_t)p->G * p->n;
    k_sketch_keys<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, (uint64_t*)d_keys);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("sketch_keys launch", e);
    e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail("sketch_keys", e);
}

// the sketch keys of genomes [g0, g0 + count) only (a 2e9-slot database is generated and handed over chunk by chunk)
__global__ void k_sketch_keys_range(syn_params p, uint32_t g0, uint64_t total, uint64_t* keys) {
    uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (s >= total) return;
    uint64_t hi, lo;
    syn_sketch_key(&p, g0 + (uint32_t)(s / p.n), (uint32_t)(s % p.n), &hi, &lo);
    keys[2 * s] = hi; keys[2 * s + 1] = lo;
}
SYN_API int syn_cuda_gen_sketch_keys_range(const syn_params* p, uint32_t g0, uint32_t count, void* d_keys, void* stream) {
    uint64_t total = (uint64_t)count * p->n;
    k_sketch_keys_range<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*p, g0, total, (uint64_t*)d_keys);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("sketch_keys_range launch", e);
    e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail("sketch_keys_range", e);
}

// bases: packed_sizes(nreads, L)[0] bytes, nmask: packed_sizes(...)[1] bytes (both multiples of 16, device)
SYN_API int syn_cuda_gen_reads_packed(const syn_params* p, uint64_t r0, uint64_t nreads, void* d_bases, void* d_nmask,
                                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    uint64_t* cum_h = (uint64_t*)malloc(sizeof(uint64_t) * p->n_present);
    uint64_t acc = 0;
    for (uint32_t i = 0; i < p->n_present; ++i) { acc += syn_present_weight(p, i); cum_h[i] = acc; }
    uint64_t* cum_d = nullptr;
    cudaError_t e = cudaMalloc(&cum_d, sizeof(uint64_t) * p->n_present);
    if (e != cudaSuccess) { free(cum_h); return fail("cudaMalloc", e); }
    cudaMemcpyAsync(cum_d, cum_h, sizeof(uint64_t) * p->n_present, cudaMemcpyHostToDevice, st);
    const uint64_t nb = nreads * p->read_len, nwords = (nb + 63) / 64;
    const uint64_t wb = ((nwords * 16 + 15) / 16 * 16) / 16, wm = ((nwords * 8 + 15) / 16 * 16) / 8;
    const uint64_t nthreads = wb > wm ? wb : wm;
    k_reads_packed<<<(unsigned)((nthreads + 127) / 128), 128, 0, st>>>(*p, cum_d, r0, nreads, wb, wm, (uint8_t*)d_bases, (uint8_t*)d_nmask);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(cum_d); free(cum_h);
    return e == cudaSuccess ? 0 : fail("reads_packed", e);
}
