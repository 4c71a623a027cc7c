// Host build of metalign_b200/csrc/kmer.cuh for tests/test_kmer_math.py (the header is __host__ __device__).
#include "../metalign_b200/csrc/kmer.cuh"
extern "C" {
void t_rc(unsigned long long hi, unsigned long long lo, unsigned k, unsigned long long* out) { key128 a{hi, lo}; key128 r = key_rc(a, k); out[0] = r.hi; out[1] = r.lo; }
void t_canon(unsigned long long hi, unsigned long long lo, unsigned k, unsigned long long* out) { key128 a{hi, lo}; key128 r = key_canon(a, k); out[0] = r.hi; out[1] = r.lo; }
void t_sub(unsigned long long hi, unsigned long long lo, unsigned K, unsigned off, unsigned k, unsigned long long* out) { key128 a{hi, lo}; key128 r = key_sub(a, K, off, k); out[0] = r.hi; out[1] = r.lo; }
void t_prefix(unsigned long long hi, unsigned long long lo, unsigned K, unsigned k, unsigned long long* out) { key128 a{hi, lo}; key128 r = key_prefix(a, K, k); out[0] = r.hi; out[1] = r.lo; }
void t_shl(unsigned long long hi, unsigned long long lo, unsigned s, unsigned long long* out) { key128 a{hi, lo}; key128 r = key_shl(a, s); out[0] = r.hi; out[1] = r.lo; }
unsigned long long t_hash(unsigned long long hi, unsigned long long lo, unsigned K) { key128 a{hi, lo}; return key_hash(a, K); }
unsigned long long t_bucket(unsigned long long h, unsigned bbits) { return hash_bucket(h, bbits); }
unsigned t_fword(unsigned long long h, unsigned nfw) { return filter_word(h, nfw); }
unsigned t_fbit(unsigned long long h) { return filter_bit(h); }
unsigned t_fmask(unsigned long long h, unsigned fk) { return filter_mask(h, fk); }
unsigned t_fp(unsigned long long h) { return hash_fp(h); }
unsigned t_minimizer(unsigned long long hi, unsigned long long lo, unsigned K) { key128 a{hi, lo}; return key_minimizer(a, K); }
unsigned long long t_hash_sk(unsigned long long hi, unsigned long long lo, unsigned K, unsigned bbits) { key128 a{hi, lo}; return key_hash_sk(a, K, bbits); }
unsigned t_mmer_mix(unsigned f, unsigned r) { return mmer_mix(f, r); }
unsigned t_rev2_32(unsigned x) { return rev2_32h(x); }
unsigned t_mz_order(unsigned a, unsigned b) { return mz_order(a, b); }
unsigned long long t_mz_ident(unsigned a, unsigned b) { return mz_ident(a, b); }
void t_key_mz(unsigned long long hi, unsigned long long lo, unsigned K, unsigned long long* out) { key128 a{hi, lo}; key_mz(a, K, out, out + 1); }
unsigned long long t_mz_bit_index(unsigned long long z, unsigned fbits) { return mz_bit_index(z, fbits); }
unsigned t_mz_bucket(unsigned long long z, unsigned bbits) { return mz_bucket(z, bbits); }
unsigned t_mz_bit2(unsigned zhi) { return mz_bit2(zhi); }
}
