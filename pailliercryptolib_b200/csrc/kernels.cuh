// kernels.cuh -- the batch kernels of the B200 Paillier hot path.
//
// Work decomposition (all kernels): one big integer = one group of T lanes,
// 32/T groups per warp, persistent grid: group g handles elements g, g+G,
// g+2G, ... so that consecutive groups touch consecutive elements (coalesced
// 128-bit loads/stores of the batch) and every warp runs the same trip count.
// Per-group scratch (the window table) lives in a workspace slot indexed by the
// group id, so it is sized by the number of resident groups -- it stays in L2 --
// not by the batch.
#pragma once
#include <cstdint>

#include "kernel_hensel_decrypt.cuh"
#include "kernels_common.cuh"
#include "mont_core.cuh"
#include "mont_hensel.cuh"

namespace ipclb200 {

#ifndef DECRYPT_MIN_BLOCKS
#define DECRYPT_MIN_BLOCKS 3
#endif
constexpr int kMaxWindow = 6;

// Per-modulus constants in device memory, each an array of L words.
struct ModConst {
  const uint32_t* n;    // the modulus
  const uint32_t* rr;   // R^2 mod n
  const uint32_t* r3;   // R^3 mod n
  const uint32_t* one;  // R mod n   (Montgomery form of 1)
  uint32_t n0inv;       // -n^-1 mod 2^32
  uint32_t small_mod;   // n < R/4: canonicalise through two multiplies
};

// window `k` (w bits wide) of an exponent of `ew` words
__device__ __forceinline__ uint32_t exp_window(const uint32_t* __restrict__ e,
                                               int ew, int k, int w) {
  int bit = k * w;
  int wi = bit >> 5, sh = bit & 31;
  uint32_t lo = (wi < ew) ? __ldg(e + wi) : 0u;
  uint32_t hi = (wi + 1 < ew) ? __ldg(e + wi + 1) : 0u;
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> sh) & ((1u << w) - 1u);
}

// Fixed-window exponentiation of one group.  xm: base in Montgomery form.
// Leaves acc = base^e in Montgomery form (almost reduced).  tab: this group's
// (1<<w)*L-word scratch.  Table loads for the next window are issued before
// the w squarings so their latency is covered.
template <int K, int T>
__device__ __forceinline__ void modexp_core(
    uint32_t (&acc)[K], const uint32_t (&xm)[K], const uint32_t (&n)[K],
    uint32_t n0inv, const uint32_t* __restrict__ one,
    const uint32_t* __restrict__ e, int ew, int ebits, int w,
    uint32_t* __restrict__ tab) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  {
    uint32_t t[K];
    M::load(t, one);
    M::store(tab, t);
    M::store(tab + L, xm);
#pragma unroll
    for (int j = 0; j < K; j++) t[j] = xm[j];
    for (int i = 2; i < (1 << w); i++) {
      M::mul(t, t, xm, n, n0inv);
      M::store(tab + (size_t)i * L, t);
    }
  }
  int nwin = (ebits + w - 1) / w;
  if (nwin < 1) nwin = 1;
  M::load(acc, tab + (size_t)exp_window(e, ew, nwin - 1, w) * L);
  for (int k = nwin - 2; k >= 0; k--) {
    uint32_t t[K];
    M::load(t, tab + (size_t)exp_window(e, ew, k, w) * L);
    // w squarings then the table multiply, through ONE inlined copy of the
    // multiply so the hot loop stays inside the instruction cache
#pragma unroll 1
    for (int s = 0; s <= w; s++) {
      uint32_t b[K];
#pragma unroll
      for (int j = 0; j < K; j++) b[j] = (s == w) ? t[j] : acc[j];
      M::mul(acc, acc, b, n, n0inv);
    }
  }
}

// Exponentiation with a SHARED exponent (p-1, q-1, lambda, n): every group
// follows the same host-precomputed sliding-window schedule, so there is no
// exponent scanning on the device and no divergence.  sched[0] = number of odd
// powers in the table (x, x^3, ..., 2^(w-1) entries); sched[1] = table index of
// the leading window; then one byte per step: 0 = square, k > 0 = multiply by
// odd power number k-1; 0xff terminates.  Against the fixed window this saves
// ~4 % of the products and halves the table (16 entries at w = 5), which keeps
// the per-group scratch L2-resident.
template <int K, int T>
__device__ __forceinline__ void modexp_sched_core(
    uint32_t (&acc)[K], const uint32_t (&xm)[K], const uint32_t (&n)[K],
    uint32_t n0inv, const uint8_t* __restrict__ sched,
    uint32_t* __restrict__ tab) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  const int nodd = sched[0];
  {
    uint32_t t[K], x2[K];
    M::store(tab, xm);
    M::mul(x2, xm, xm, n, n0inv);
#pragma unroll
    for (int j = 0; j < K; j++) t[j] = xm[j];
    for (int i = 1; i < nodd; i++) {
      M::mul(t, t, x2, n, n0inv);
      M::store(tab + (size_t)i * L, t);
    }
  }
  M::load(acc, tab + (size_t)sched[1] * L);
  const uint8_t* op = sched + 2;
#pragma unroll 1
  for (uint32_t o = __ldg(op); o != 0xffu; o = __ldg(++op)) {
    uint32_t b[K];
    if (o) {
      M::load(b, tab + (size_t)(o - 1) * L);
    } else {
#pragma unroll
      for (int j = 0; j < K; j++) b[j] = acc[j];
    }
    M::mul(acc, acc, b, n, n0inv);
  }
}

// x (< R, any residue class) -> canonical x mod n
template <int K, int T>
__device__ __forceinline__ void canonicalize(uint32_t (&x)[K],
                                             const uint32_t (&n)[K],
                                             uint32_t n0inv,
                                             const uint32_t* __restrict__ rr,
                                             uint32_t small_mod) {
  using M = Mont<K, T>;
  if (small_mod) {
    uint32_t t[K];
    M::load(t, rr);
    M::mul(x, x, t, n, n0inv);
    M::from_mont(x, x, n, n0inv);
  } else {
#pragma unroll 1
    for (int i = 0; i < 3; i++) M::sub_n_if_ge(x, n);
  }
}

// --------------------------------------------------------------------------
// K1: generic batched modexp  (ipcl::ippModExp, ipcl/mod_exp.cpp:655-678)
// --------------------------------------------------------------------------
struct ModexpParams {
  const uint32_t* base;
  size_t base_stride;  // words, 0 = shared
  const uint32_t* exp;
  size_t exp_stride;  // words, 0 = shared
  int exp_words;
  int exp_bits;
  // modulus constants: shared (stride 0) or per element (stride L / 1)
  const uint32_t* n;
  const uint32_t* rr;
  const uint32_t* one;
  const uint32_t* n0inv;
  size_t mod_stride;
  size_t n0_stride;
  uint32_t* out;
  size_t count;
  uint32_t* table_ws;
  int window;
  unsigned int* work_counter;
  // shared exponent: host-built sliding-window schedule (modexp_sched_core),
  // nullptr = scan the exponent with the fixed window
  const uint8_t* sched;
};

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    modexp_kernel(const ModexpParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* tab = p.table_ws + gid * ((size_t)L << p.window);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= nchunks) break;
    const size_t inst = (size_t)w * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    uint32_t n[K], x[K], acc[K];
    M::load(n, p.n + ii * p.mod_stride);
    const uint32_t n0inv = __ldg(p.n0inv + ii * p.n0_stride);
    {
      uint32_t rr[K];
      M::load(rr, p.rr + ii * p.mod_stride);
      M::load(x, p.base + ii * p.base_stride);
      M::mul(x, x, rr, n, n0inv);
    }
    if (p.sched)
      modexp_sched_core<K, T>(acc, x, n, n0inv, p.sched, tab);
    else
      modexp_core<K, T>(acc, x, n, n0inv, p.one + ii * p.mod_stride,
                        p.exp + ii * p.exp_stride, p.exp_words, p.exp_bits,
                        p.window, tab);
    M::from_mont(x, acc, n, n0inv);
    if (valid) M::store(p.out + inst * L, x);
  }
}

// --------------------------------------------------------------------------
// K2: batched modular multiply  (CipherText::raw_add, ciphertext.cpp:135-141)
// --------------------------------------------------------------------------
struct ModmulParams {
  const uint32_t* a;
  const uint32_t* b;
  size_t b_stride;  // 0 = shared
  ModConst m;
  uint32_t* out;
  size_t count;
  unsigned int* work_counter;
};

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    modmul_kernel(const ModmulParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  uint32_t n[K], rr[K];
  M::load(n, p.m.n);
  M::load(rr, p.m.rr);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= nchunks) break;
    const size_t inst = (size_t)w * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    uint32_t a[K], b[K];
    M::load(a, p.a + ii * L);
    M::load(b, p.b + ii * p.b_stride);
    M::mul(a, a, rr, n, p.m.n0inv);  // a*R
    M::mul(a, a, b, n, p.m.n0inv);   // a*b  (< R)
    canonicalize<K, T>(a, n, p.m.n0inv, p.m.rr, p.m.small_mod);
    if (valid) M::store(p.out + inst * L, a);
  }
}

// --------------------------------------------------------------------------
// K3: fused Paillier encrypt  (PublicKey::raw_encrypt + applyObfuscator,
//     ipcl/pub_key.cpp:51-110), modulus n^2 of L = 2*NL words.
//   gm  = n*pt + 1            (one multiply by n*R mod n^2, then +1)
//   obf = hs^r  (DJN: fixed-base comb table, no squarings -- K5)
//       | r^n   (non-DJN: fixed-window modexp, shared exponent n)
//   ct  = obf * gm mod n^2    (obf stays in Montgomery form, so this multiply
//                              also leaves Montgomery form)
// --------------------------------------------------------------------------
struct EncryptParams {
  const uint32_t* pt;
  int pt_words;
  const uint32_t* r;
  int r_words;
  int r_bits;
  ModConst m;           // n^2
  const uint32_t* nR;   // n * R mod n^2
  const uint32_t* n_exp;  // n as exponent (non-DJN), NL words
  int n_exp_words;
  const uint8_t* sched_n;  // sliding-window schedule of n (non-DJN) or nullptr
  const uint32_t* hs_m;  // hs in Montgomery form (DJN generic path), L words
  const uint32_t* comb;  // comb table [nwin][1<<cw][L] or nullptr
  int comb_w;
  int comb_windows;
  int mode;  // 0: no obfuscator, 1: DJN comb, 2: DJN generic window, 3: non-DJN
  uint32_t* ct;
  size_t count;
  uint32_t* table_ws;
  int window;
  unsigned int* work_counter;
};

// load `words` words (zero extended to L) spread over the group
template <int K, int T>
__device__ __forceinline__ void load_padded(uint32_t (&x)[K],
                                            const uint32_t* __restrict__ p,
                                            int words) {
  const int base = Mont<K, T>::lane_t() * K;
#pragma unroll
  for (int j = 0; j < K; j++) x[j] = (base + j < words) ? p[base + j] : 0u;
}

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    encrypt_kernel(const EncryptParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* tab = p.table_ws + gid * ((size_t)L << p.window);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  uint32_t n[K];
  M::load(n, p.m.n);
  const uint32_t n0inv = p.m.n0inv;
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= nchunks) break;
    const size_t inst = (size_t)w * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    uint32_t gm[K];
    {
      uint32_t t[K];
      load_padded<K, T>(gm, p.pt + ii * (size_t)p.pt_words, p.pt_words);
      M::load(t, p.nR);
      M::mul(gm, gm, t, n, n0inv);  // n*pt mod n^2 (< R)
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = 0;
      uint32_t c = M::group_add(gm, t, 1u);  // + 1
      if (__any_sync(IPCLB200_FULL_MASK, c)) M::cond_sub_n(gm, n, c);
    }
    if (p.mode == 0) {
      canonicalize<K, T>(gm, n, n0inv, p.m.rr, p.m.small_mod);
      if (valid) M::store(p.ct + inst * L, gm);
      continue;
    }
    uint32_t acc[K];
    if (p.mode == 1) {
      const uint32_t* e = p.r + ii * (size_t)p.r_words;
      const size_t wstride = (size_t)L << p.comb_w;
      M::load(acc, p.comb + (size_t)exp_window(e, p.r_words, 0, p.comb_w) * L);
      uint32_t t[K];
      if (p.comb_windows > 1)
        M::load(t, p.comb + wstride +
                       (size_t)exp_window(e, p.r_words, 1, p.comb_w) * L);
      for (int k = 1; k < p.comb_windows; k++) {
        uint32_t tn[K];
        if (k + 1 < p.comb_windows)
          M::load(tn, p.comb + (size_t)(k + 1) * wstride +
                          (size_t)exp_window(e, p.r_words, k + 1, p.comb_w) * L);
        M::mul(acc, acc, t, n, n0inv);
#pragma unroll
        for (int j = 0; j < K; j++) t[j] = tn[j];
      }
    } else if (p.mode == 2) {
      uint32_t x[K];
      M::load(x, p.hs_m);
      modexp_core<K, T>(acc, x, n, n0inv, p.m.one,
                        p.r + ii * (size_t)p.r_words, p.r_words, p.r_bits,
                        p.window, tab);
    } else {
      uint32_t x[K], t[K];
      load_padded<K, T>(x, p.r + ii * (size_t)p.r_words, p.r_words);
      M::load(t, p.m.rr);
      M::mul(x, x, t, n, n0inv);
      if (p.sched_n)
        modexp_sched_core<K, T>(acc, x, n, n0inv, p.sched_n, tab);
      else
        modexp_core<K, T>(acc, x, n, n0inv, p.m.one, p.n_exp, p.n_exp_words,
                          p.n_exp_words * 32, p.window, tab);
    }
    M::mul(acc, acc, gm, n, n0inv);  // obf*gm, out of Montgomery form
    canonicalize<K, T>(acc, n, n0inv, p.m.rr, p.m.small_mod);
    if (valid) M::store(p.ct + inst * L, acc);
  }
}

// --------------------------------------------------------------------------
// K5: fixed-base comb table for the DJN obfuscator hs^r (pub_key.cpp:51-64:
//     the same base hs for every element).  comb[i][j] = hs^(j * 2^(w*i)) in
//     Montgomery form.  Phase A (one group): the spine g_i = hs^(2^(w*i)).
//     Phase B (one group per window): the 2^w multiples of g_i.
// --------------------------------------------------------------------------
struct CombParams {
  ModConst m;
  const uint32_t* hs_m;  // hs in Montgomery form
  uint32_t* comb;
  int w;     // window width, entries per window = 2^w
  int w_lo;  // two-level build: entry j = (b << w_lo) | a is row[b << w_lo] * row[a]
  int windows;
};

template <int K, int T>
__global__ void __launch_bounds__(32) comb_spine_kernel(const CombParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  // every group of the single warp computes the same spine; group 0 stores
  uint32_t n[K], g[K];
  M::load(n, p.m.n);
  M::load(g, p.hs_m);
  const bool writer = (threadIdx.x / T) == 0;
  const size_t wstride = (size_t)L << p.w;
  for (int i = 0; i < p.windows; i++) {
    if (writer) M::store(p.comb + (size_t)i * wstride + L, g);
    for (int s = 0; s < p.w; s++) {
      M::mul(g, g, g, n, p.m.n0inv);
      // g_i^(2^w_lo): the generator of the upper chain of window i
      if (writer && s + 1 == p.w_lo && p.w_lo < p.w)
        M::store(p.comb + (size_t)i * wstride + ((size_t)L << p.w_lo), g);
    }
  }
}

// level 1: the two sequential chains of each window, one group per chain:
//   lower  row[a]         = g^a            a < 2^w_lo   (row[1] = g given)
//   upper  row[b << w_lo] = (g^(2^w_lo))^b b < 2^(w - w_lo)
template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    comb_fill_kernel(const CombParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  const int gpb = blockDim.x / T;
  const int gid = blockIdx.x * gpb + threadIdx.x / T;
  const bool two = p.w_lo < p.w;
  const int nchains = two ? 2 * p.windows : p.windows;
  const bool valid = gid < nchains;
  const int ch = valid ? gid : nchains - 1;
  const int i = two ? ch >> 1 : ch;
  const bool upper = two && (ch & 1);
  const size_t step = upper ? ((size_t)1 << p.w_lo) : 1;
  const int count = upper ? (1 << (p.w - p.w_lo)) : (1 << p.w_lo);
  uint32_t* row = p.comb + (size_t)i * ((size_t)L << p.w);
  uint32_t n[K], g[K], t[K];
  M::load(n, p.m.n);
  M::load(g, row + step * L);
  if (!upper) {
    M::load(t, p.m.one);
    if (valid) M::store(row, t);
  }
#pragma unroll
  for (int j = 0; j < K; j++) t[j] = g[j];
  for (int j = 2; j < count; j++) {
    M::mul(t, t, g, n, p.m.n0inv);
    if (valid) M::store(row + (size_t)j * step * L, t);
  }
}

// level 2: every remaining entry is one product of a lower and an upper entry
template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    comb_expand_kernel(const CombParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  const size_t ngroups = (size_t)gridDim.x * gpb;
  const size_t per_win = (size_t)1 << p.w;
  const size_t total = per_win * (size_t)p.windows;
  const size_t lo_mask = ((size_t)1 << p.w_lo) - 1;
  uint32_t n[K];
  M::load(n, p.m.n);
  const size_t iters = (total + ngroups - 1) / ngroups;
  for (size_t it = 0; it < iters; it++) {
    const size_t e = it * ngroups + gid;
    const bool in = e < total;
    const size_t ee = in ? e : total - 1;
    const size_t j = ee & (per_win - 1);
    uint32_t* row = p.comb + (ee >> p.w) * (per_win * L);
    const bool need = in && (j & lo_mask) != 0 && (j & ~lo_mask) != 0;
    uint32_t a[K], b[K];
    M::load(a, row + (j & lo_mask) * L);
    M::load(b, row + (j & ~lo_mask) * L);
    M::mul(a, a, b, n, p.m.n0inv);
    if (need) M::store(row + j * L, a);
  }
}

// --------------------------------------------------------------------------
// K4: fused CRT decrypt, modexp part  (PrivateKey::decryptCRT,
//     ipcl/pri_key.cpp:114-146).  Two tasks per ciphertext (mod p^2, mod q^2),
//     L = words of p^2.  Prologue: ct mod p^2 by one Montgomery reduction of
//     the low half plus the high half; then x = ct^(p-1) mod p^2, canonical.
//     The L-function / hp / CRT recombination runs in crt_finish_kernel.
// --------------------------------------------------------------------------
struct DecryptCrtParams {
  const uint32_t* ct;  // count x 2L words
  ModConst m0, m1;     // p^2, q^2 (no arrays: a runtime-indexed kernel
                       // parameter would be copied to local memory)
  const uint8_t *sched0, *sched1;  // sliding-window schedules of p-1, q-1
  uint32_t* x;  // out: count x 2 x L words
  size_t count;
  uint32_t* table_ws;
  int table_entries;  // odd powers per table
  unsigned int* work_counter;
};

// the integer-pipe role: claims chunks until the batch is exhausted
template <int K, int T>
__device__ __forceinline__ void decrypt_int_role(const DecryptCrtParams& p,
                                                 size_t gid) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;  // groups per warp
  uint32_t* tab = p.table_ws + gid * ((size_t)L * p.table_entries);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    // a chunk = the GW ciphertexts of one warp on ONE side (p^2 or q^2), so
    // all groups of the warp follow the same schedule
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= 2u * nchunks) break;
    const int side = (int)(w & 1u);
    const size_t inst = (size_t)(w >> 1) * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* mn = side ? p.m1.n : p.m0.n;
    const uint32_t* mr3 = side ? p.m1.r3 : p.m0.r3;
    const uint32_t n0inv = side ? p.m1.n0inv : p.m0.n0inv;
    uint32_t n[K];
    M::load(n, mn);
    const uint32_t* c = p.ct + ii * (size_t)(2 * L);
    uint32_t x[K], acc[K];
    {
      uint32_t lo[K], hi[K], t[K];
      M::load(lo, c);
      M::load(hi, c + L);
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = 0;
      if (M::lane_t() == 0) t[0] = 1;
      M::mul(lo, lo, t, n, n0inv);  // lo * R^-1   (<= n)
      uint32_t cy = M::group_add(lo, hi, 0u);  // ct * R^-1 mod n, < R + n
      if (__any_sync(IPCLB200_FULL_MASK, cy)) M::cond_sub_n(lo, n, cy);
      M::load(t, mr3);
      M::mul(x, lo, t, n, n0inv);  // ct * R mod n: Montgomery form
    }
    modexp_sched_core<K, T>(acc, x, n, n0inv, side ? p.sched1 : p.sched0, tab);
    M::from_mont(x, acc, n, n0inv);
    if (valid) M::store(p.x + (inst * 2 + side) * L, x);
  }
}

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads, DECRYPT_MIN_BLOCKS)
    decrypt_crt_kernel(const DecryptCrtParams p) {
  const size_t gpb = blockDim.x / T;
  decrypt_int_role<K, T>(p, blockIdx.x * gpb + threadIdx.x / T);
}


// K4h (decrypt_hensel_kernel) lives in kernel_hensel_decrypt.cuh, included above

// --------------------------------------------------------------------------
// K3h: DJN encrypt in two-digit arithmetic mod n^2 (digits base n, LH = words of
//     n), mont_hensel.cuh.  obf = hs^r is the product of one fixed-base table
//     entry per window (K5), entries stored as pairs: 5 half-width limb products
//     per window instead of 8.  Window 0 is stored in plain form, the others in
//     Montgomery form, so the running product stays plain: obf = u0 - uw*n.
//     With a = u0 mod n (u0 = a + c*n):
//         ct = obf * (1 + m*n) = a + ((c + a*m - uw) mod n) * n      (mod n^2)
//     -- canonical by construction: one half-width modular product a*m, one
//     plain LH x LH product t*n.  Entries are fetched into shared memory by TMA
//     bulk copies, double buffered, one window ahead.
// --------------------------------------------------------------------------
struct EncryptHenselParams {
  const uint32_t* pt;
  int pt_words;
  const uint32_t* r;
  int r_words;
  const uint32_t* blk;   // n | R^2 mod n            (2*LH words)
  uint32_t n0inv;
  const uint32_t* comb;  // [windows][1 << w][2*LH]: pairs
  int comb_w;
  int comb_windows;      // windows covering this batch's exponents
  uint32_t* ct;          // count x 2*LH words
  size_t count;
  unsigned int* work_counter;
};

// group area: the HMont layout (S0 | S1 | SQ | entry buffer 0) + entry buffer 1
template <int K, int T>
constexpr size_t encrypt_hensel_smem_bytes(int threads) {
  return (size_t)(threads / T) * (7 * K * T + 4) * sizeof(uint32_t) +
         (size_t)(threads / 32) * 2 * sizeof(uint64_t);
}

template <int K, int T, int MINB, int ROWS>
__global__ void __launch_bounds__(kBlockThreads, MINB)
    encrypt_hensel_kernel(const EncryptHenselParams p) {
  using M = Mont<K, T>;
  using H = HMont<K, T, ROWS>;
  constexpr int LH = K * T;
  constexpr int GW = 32 / T;
  constexpr int kStride = 7 * LH + 4;
  constexpr int kBuf = 2 * LH;
  extern __shared__ __align__(16) uint32_t hensel_smem[];
  const int gib = threadIdx.x / T;
  uint32_t* sm = hensel_smem + (size_t)gib * kStride;
  uint32_t* sS0 = sm + H::kS0;
  uint32_t* sSQ = sm + H::kSQ;
  uint32_t* sBuf = sm + H::kT0;  // two entry buffers of 2*LH words
  uint64_t* bar = reinterpret_cast<uint64_t*>(hensel_smem +
                                              (size_t)(blockDim.x / T) * kStride) +
                  2 * (threadIdx.x >> 5);
  const bool warp_leader = (threadIdx.x & 31) == 0;
  const bool group_leader = M::lane_t() == 0;
  if (warp_leader) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase[2] = {0, 0};
  constexpr uint32_t kEntryBytes = 2 * LH * sizeof(uint32_t);
  const size_t wstride = (size_t)(2 * LH) << p.comb_w;
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  uint32_t n[K];
  M::load(n, p.blk);
  const uint32_t n0inv = p.n0inv;
  for (;;) {
    const unsigned int wk = claim_chunk(p.work_counter);
    if (wk >= nchunks) break;
    const size_t inst = (size_t)wk * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* e = p.r + ii * (size_t)p.r_words;
    uint32_t x0[K], w[K];
    // window 0 (plain form) straight into registers, window 1 on its way
    {
      const uint32_t* ent = p.comb + (size_t)exp_window(e, p.r_words, 0, p.comb_w) * kBuf;
      M::load(x0, ent);
      M::load(w, ent + LH);
    }
    __syncwarp();
    fence_async_proxy();
    if (p.comb_windows > 1) {
      if (warp_leader) mbar_expect_tx(bar, GW * kEntryBytes);
      __syncwarp();
      if (group_leader)
        bulk_g2s(sBuf, p.comb + wstride + (size_t)exp_window(e, p.r_words, 1, p.comb_w) * kBuf,
                 kEntryBytes, bar);
    }
#pragma unroll 1
    for (int k = 1; k < p.comb_windows; k++) {
      const int b = (k - 1) & 1;
      if (k + 1 < p.comb_windows) {
        // the other buffer was last read by the multiply of window k-1
        __syncwarp();
        fence_async_proxy();
        if (warp_leader) mbar_expect_tx(bar + (b ^ 1), GW * kEntryBytes);
        __syncwarp();
        if (group_leader)
          bulk_g2s(sBuf + (b ^ 1) * kBuf,
                   p.comb + (size_t)(k + 1) * wstride +
                       (size_t)exp_window(e, p.r_words, k + 1, p.comb_w) * kBuf,
                   kEntryBytes, bar + (b ^ 1));
      }
      mbar_wait(bar + b, phase[b]);
      phase[b] ^= 1u;
      H::mul(x0, w, sBuf + b * kBuf, sm, n, n0inv);
    }
    // ---- am = a * m mod n: two half-width Montgomery products (m, then R^2).
    // a*m only matters mod n, so the not yet reduced u0 (< R) serves as a ----------
    uint32_t t[K];
    {
      uint32_t m[K];
      load_padded<K, T>(m, p.pt + ii * (size_t)p.pt_words, p.pt_words);
      __syncwarp();
      H::put(sS0, m);
      __syncwarp();
      H::pass_a(t, x0, sS0, sSQ, n, n0inv);  // u0*m/R
      M::load(m, p.blk + LH);
      __syncwarp();
      H::put(sS0, m);
      __syncwarp();
      uint32_t t2[K];
      H::pass_a(t2, t, sS0, sSQ, n, n0inv);  // u0*m mod n, < R
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = t2[j];
    }
    M::sub_n_if_ge(t, n);
    // ---- ct = a + ((am + c - uw) mod n) * n ------------------------------------------
    uint32_t hi[K], lo[K];
    H::to_canonical(lo, hi, x0, w, t, sm, n);
    if (valid) {
      M::store(p.ct + inst * (size_t)(2 * LH), lo);
      M::store(p.ct + inst * (size_t)(2 * LH) + LH, hi);
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------
// K1h: batched modexp mod n^2 in two-digit arithmetic, for a modulus that is
//     the square of an n that fills its LH words (CipherText * PlainText,
//     ipcl/ciphertext.cpp:143-162: a^b mod n^2 with per-element b).  Prologue:
//     the 2*LH-word base -> Montgomery pair; fixed-window ladder over a table of
//     all 2^w powers in the group's workspace slot (window entries fetched by TMA
//     bulk copies while the squarings run); epilogue: leave Montgomery form with
//     one product by the pair of 1, then pair -> canonical residue.
// --------------------------------------------------------------------------
struct ModexpHenselParams {
  const uint32_t* base;  // count x 2*LH words (any value < R^2)
  const uint32_t* exp;
  size_t exp_stride;     // words, 0 = shared
  int exp_words;
  int exp_bits;
  int window;
  const uint32_t* blk;   // n | pairs of R, R^2 (Montgomery form of chunk weights 1, R)
  uint32_t n0inv;
  uint32_t* out;         // count x 2*LH words
  size_t count;
  uint32_t* table_ws;    // per group (1 << window) x 2*LH words
  unsigned int* work_counter;
};

template <int K, int T, int MINB, int ROWS>
__global__ void __launch_bounds__(kBlockThreads, MINB)
    modexp_hensel_kernel(const ModexpHenselParams p) {
  using M = Mont<K, T>;
  using H = HMont<K, T, ROWS>;
  constexpr int LH = K * T;
  constexpr int GW = 32 / T;
  extern __shared__ __align__(16) uint32_t hensel_smem[];
  const int gib = threadIdx.x / T;
  uint32_t* sm = hensel_smem + (size_t)gib * H::kStride;
  uint64_t* bar = reinterpret_cast<uint64_t*>(hensel_smem +
                                              (size_t)(blockDim.x / T) * H::kStride) +
                  (threadIdx.x >> 5);
  const bool warp_leader = (threadIdx.x & 31) == 0;
  const bool group_leader = M::lane_t() == 0;
  if (warp_leader) mbar_init(bar, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t phase = 0;
  const size_t gid = (size_t)blockIdx.x * (blockDim.x / T) + gib;
  const int w = p.window;
  uint32_t* tab = p.table_ws + gid * ((size_t)(2 * LH) << w);
  constexpr uint32_t kEntryBytes = 2 * LH * sizeof(uint32_t);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  uint32_t n[K];
  M::load(n, p.blk);
  const uint32_t n0inv = p.n0inv;
  int nwin = (p.exp_bits + w - 1) / w;
  if (nwin < 1) nwin = 1;
  for (;;) {
    const unsigned int wk = claim_chunk(p.work_counter);
    if (wk >= nchunks) break;
    const size_t inst = (size_t)wk * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* e = p.exp + ii * p.exp_stride;
    uint32_t x0[K], xw[K];
    H::enter(x0, xw, p.base + ii * (size_t)(2 * LH), 2, p.blk + LH, sm, n, n0inv);
    // ---- table: x^0 = 1 as the pair (R - n, n - 1), x^1, x^i = x^(i-1) * x ---------
    {
      uint32_t o0[K], o1[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        o0[j] = ~n[j];
        o1[j] = n[j];
      }
      if (M::lane_t() == 0) {
        o0[0] += 1u;
        o1[0] -= 1u;
      }
      M::store(tab, o0);
      M::store(tab + LH, o1);
      M::store(tab + 2 * LH, x0);
      M::store(tab + 3 * LH, xw);
      __syncwarp();
      H::put(sm + H::kT0, x0);
      H::put(sm + H::kT0 + LH, xw);
      __syncwarp();
#pragma unroll 1
      for (int i = 2; i < (1 << w); i++) {
        H::step(x0, xw, true, sm + H::kT0, sm, n, n0inv);
        M::store(tab + (size_t)i * 2 * LH, x0);
        M::store(tab + (size_t)i * 2 * LH + LH, xw);
      }
    }
    // ---- ladder -------------------------------------------------------------------
    {
      const uint32_t first = exp_window(e, p.exp_words, nwin - 1, w);
      M::load(x0, tab + (size_t)first * 2 * LH);
      M::load(xw, tab + (size_t)first * 2 * LH + LH);
    }
#pragma unroll 1
    for (int k = nwin - 2; k >= 0; k--) {
      __syncwarp();
      fence_async_proxy();
      if (warp_leader) mbar_expect_tx(bar, GW * kEntryBytes);
      __syncwarp();
      if (group_leader)
        bulk_g2s(sm + H::kT0, tab + (size_t)exp_window(e, p.exp_words, k, w) * 2 * LH,
                 kEntryBytes, bar);
#pragma unroll 1
      for (int sq = 0; sq <= w; sq++) {
        const bool is_mul = sq == w;
        if (is_mul) {
          mbar_wait(bar, phase);
          phase ^= 1u;
        }
        H::step(x0, xw, is_mul, sm + H::kT0, sm, n, n0inv);
      }
    }
    // ---- leave Montgomery form: times the raw pair (1, 0), then canonical -------------
    {
      uint32_t one[K], zero[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        one[j] = 0;
        zero[j] = 0;
      }
      if (M::lane_t() == 0) one[0] = 1;
      __syncwarp();
      H::put(sm + H::kT0, one);
      H::put(sm + H::kT0 + LH, zero);
      __syncwarp();
      H::step(x0, xw, true, sm + H::kT0, sm, n, n0inv);
      uint32_t lo[K], hi[K];
      H::to_canonical(lo, hi, x0, xw, zero, sm, n);
      if (valid) {
        M::store(p.out + inst * (size_t)(2 * LH), lo);
        M::store(p.out + inst * (size_t)(2 * LH) + LH, hi);
      }
    }
    __syncwarp();
  }
}

// K5h: full-width Montgomery table entries -> pairs.  in[i] = v_i * R^2 mod n^2
// (R = 2^(32 LH), the full-width radix is R^2), 2*LH words, any value < R^2.
// out[i] = the pair of v_i * R mod n^2 (Montgomery form; windows >= 1) or of v_i
// (plain; window 0): sum of the two LH-word chunks times the constants
// R^(j-1) resp. R^(j-2) as raw pairs (HMont::enter).
struct CombPairsParams {
  const uint32_t* in;
  uint32_t* out;
  size_t count;          // entries
  size_t plain_entries;  // the first entries (window 0) go to plain form
  const uint32_t* blk;   // n (LH)
  const uint32_t* consts_mont;   // 2 pairs
  const uint32_t* consts_plain;  // 2 pairs
  uint32_t n0inv;
  unsigned int* work_counter;
};

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads, 3) comb_pairs_kernel(const CombPairsParams p) {
  using M = Mont<K, T>;
  using H = HMont<K, T, 4>;
  constexpr int LH = K * T;
  constexpr int GW = 32 / T;
  extern __shared__ __align__(16) uint32_t hensel_smem[];
  uint32_t* sm = hensel_smem + (size_t)(threadIdx.x / T) * H::kStride;
  uint32_t n[K];
  M::load(n, p.blk);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int wk = claim_chunk(p.work_counter);
    if (wk >= nchunks) break;
    const size_t inst = (size_t)wk * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    // all groups of a warp convert entries of the same kind (plain_entries is a
    // multiple of the chunk size: a power of two >= 16)
    const bool plain = ((size_t)wk * GW) < p.plain_entries;
    uint32_t x0[K], w[K];
    H::enter(x0, w, p.in + ii * (size_t)(2 * LH), 2, plain ? p.consts_plain : p.consts_mont, sm,
             n, p.n0inv);
    if (valid) {
      M::store(p.out + inst * (size_t)(2 * LH), x0);
      M::store(p.out + inst * (size_t)(2 * LH) + LH, w);
    }
    __syncwarp();
  }
}

// --------------------------------------------------------------------------
// CRT recombination, one thread per ciphertext (tiny next to the modexps:
// ~8k MAC32 against ~20M).  pri_key.cpp:141-157:
//   dp = L_p(xp) * hp mod p,  dq = L_q(xq) * hq mod q,  L_a(x) = (x-1)/a
//   pt = dp + ((dq - dp) * p^-1 mod q) * p
// Exact division by the odd prime is a multiplication by its inverse mod
// 2^32 limb by limb; the two modular products are word-serial Montgomery
// multiplies against constants pre-multiplied by R.
// --------------------------------------------------------------------------
constexpr int kMaxPrimeWords = 128;  // words of n for the RAW tail, 2x a prime

struct CrtFinishParams {
  const uint32_t* x;  // count x 2 x (2*PL) words: xp, xq
  const uint32_t* p;  // PL words
  const uint32_t* q;
  const uint32_t* hpR;    // hp * Rp mod p
  const uint32_t* hqR;    // hq * Rq mod q
  const uint32_t* pinvR;  // (p^-1 mod q) * Rq mod q
  uint32_t p_inv32;       // p^-1 mod 2^32   (exact division)
  uint32_t q_inv32;
  uint32_t p_n0inv;  // -p^-1 mod 2^32  (Montgomery)
  uint32_t q_n0inv;
  int pl;
  int xl;        // words per residue in x (the kernel size class of p^2)
  uint32_t* pt;  // count x 2*PL words
  size_t count;
};

// q[0..pl) = (x - 1) / d for x = 1 (mod d), x < d^2; x has 2*pl words
__device__ inline void exact_div_minus1(uint32_t* q, const uint32_t* x,
                                        const uint32_t* d, uint32_t dinv32,
                                        int pl) {
  uint32_t t[kMaxPrimeWords];
  // t = low pl words of x - 1 (the quotient only depends on them)
  uint32_t borrow = 1;
  for (int i = 0; i < pl; i++) {
    uint32_t v = x[i];
    t[i] = v - borrow;
    borrow = (v < borrow) ? 1u : 0u;
  }
  for (int i = 0; i < pl; i++) {
    uint32_t qi = t[i] * dinv32;
    q[i] = qi;
    // t -= qi * d << (32 i), only words < pl matter
    uint64_t carry = 0;
    uint32_t br = 0;
    for (int j = 0; i + j < pl; j++) {
      uint64_t pr = (uint64_t)qi * d[j] + carry;
      carry = pr >> 32;
      uint32_t sub = (uint32_t)pr;
      uint32_t v = t[i + j];
      uint32_t r1 = v - sub;
      uint32_t b1 = v < sub;
      uint32_t r2 = r1 - br;
      uint32_t b2 = r1 < br;
      t[i + j] = r2;
      br = b1 | b2;
    }
  }
}

// r = a * b * R^-1 mod m (canonical), word-serial CIOS, pl words
__device__ inline void mont_mul_serial(uint32_t* r, const uint32_t* a,
                                       const uint32_t* b, const uint32_t* m,
                                       uint32_t n0inv, int pl) {
  uint32_t t[kMaxPrimeWords + 2];
  for (int i = 0; i < pl + 2; i++) t[i] = 0;
  for (int i = 0; i < pl; i++) {
    uint64_t c = 0;
    uint32_t bi = b[i];
    for (int j = 0; j < pl; j++) {
      c += (uint64_t)a[j] * bi + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[pl];
    t[pl] = (uint32_t)c;
    t[pl + 1] = (uint32_t)(c >> 32);
    uint32_t qd = t[0] * n0inv;
    c = ((uint64_t)qd * m[0] + t[0]) >> 32;
    for (int j = 1; j < pl; j++) {
      c += (uint64_t)qd * m[j] + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[pl];
    t[pl - 1] = (uint32_t)c;
    t[pl] = t[pl + 1] + (uint32_t)(c >> 32);
  }
  // canonical: t < 2m
  bool ge = t[pl] != 0;
  if (!ge) {
    ge = true;
    for (int i = pl - 1; i >= 0; i--) {
      if (t[i] != m[i]) {
        ge = t[i] > m[i];
        break;
      }
    }
  }
  uint32_t br = 0;
  for (int i = 0; i < pl; i++) {
    uint32_t s = ge ? m[i] : 0u;
    uint32_t v = t[i];
    uint32_t r1 = v - s;
    uint32_t b1 = v < s;
    uint32_t r2 = r1 - br;
    uint32_t b2 = r1 < br;
    r[i] = r2;
    br = b1 | b2;
  }
}

// pt = dp + ((dq - dp) * p^-1 mod q) * p   (dp < p < q canonical; 2*pl words)
__device__ inline void crt_recombine(uint32_t* dst, const uint32_t* dp,
                                     const uint32_t* dq, const uint32_t* P,
                                     const uint32_t* Qm, const uint32_t* pinvR,
                                     uint32_t q_n0inv, int pl) {
  uint32_t l[kMaxPrimeWords];
  // l = (dq - dp) mod q
  uint32_t br = 0;
  for (int j = 0; j < pl; j++) {
    uint32_t v = dq[j], s = dp[j];
    uint32_t r1 = v - s;
    uint32_t b1 = v < s;
    uint32_t r2 = r1 - br;
    uint32_t b2 = r1 < br;
    l[j] = r2;
    br = b1 | b2;
  }
  if (br) {
    uint64_t c = 0;
    for (int j = 0; j < pl; j++) {
      c += (uint64_t)l[j] + Qm[j];
      l[j] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint32_t u[kMaxPrimeWords];
  mont_mul_serial(u, l, pinvR, Qm, q_n0inv, pl);
  uint32_t o[kMaxPrimeWords];
  for (int j = 0; j < 2 * pl; j++) o[j] = (j < pl) ? dp[j] : 0u;
  for (int a = 0; a < pl; a++) {
    uint64_t c = 0;
    uint32_t ua = u[a];
    for (int j = 0; j < pl; j++) {
      c += (uint64_t)ua * P[j] + o[a + j];
      o[a + j] = (uint32_t)c;
      c >>= 32;
    }
    for (int j = a + pl; c && j < 2 * pl; j++) {
      c += o[j];
      o[j] = (uint32_t)c;
      c >>= 32;
    }
  }
  for (int j = 0; j < 2 * pl; j++) dst[j] = o[j];
}

__global__ void crt_finish_kernel(const CrtFinishParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.count) return;
  const int pl = p.pl;
  const uint32_t* xp = p.x + i * (size_t)(2 * p.xl);
  const uint32_t* xq = xp + p.xl;
  uint32_t l[kMaxPrimeWords], dp[kMaxPrimeWords], dq[kMaxPrimeWords];
  exact_div_minus1(l, xp, p.p, p.p_inv32, pl);
  mont_mul_serial(dp, l, p.hpR, p.p, p.p_n0inv, pl);
  exact_div_minus1(l, xq, p.q, p.q_inv32, pl);
  mont_mul_serial(dq, l, p.hqR, p.q, p.q_n0inv, pl);
  crt_recombine(p.pt + i * (size_t)(2 * pl), dp, dq, p.p, p.q, p.pinvR, p.q_n0inv, pl);
}

// CRT recombination after decrypt_hensel_kernel, which already delivers
// mp = L_p(..)*hp mod p and mq (canonical, pl words each)
struct CrtCombineParams {
  const uint32_t* mpq;  // count x 2 x pl words
  const uint32_t* p;
  const uint32_t* q;
  const uint32_t* pinvR;
  uint32_t q_n0inv;
  int pl;
  uint32_t* pt;  // count x 2*pl words
  size_t count;
};

__global__ void crt_combine_kernel(const CrtCombineParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.count) return;
  const int pl = p.pl;
  const uint32_t* mp = p.mpq + i * (size_t)(2 * pl);
  uint32_t dp[kMaxPrimeWords], dq[kMaxPrimeWords];
  for (int j = 0; j < pl; j++) {
    dp[j] = mp[j];
    dq[j] = mp[pl + j];
  }
  crt_recombine(p.pt + i * (size_t)(2 * pl), dp, dq, p.p, p.q, p.pinvR, p.q_n0inv, pl);
}

// decryptRAW tail, one thread per ciphertext (pri_key.cpp:105-110):
//   pt = ((x - 1) / n) * mu mod n     with x = ct^lambda mod n^2
struct RawFinishParams {
  const uint32_t* x;  // count x 2*NL words
  const uint32_t* n;  // NL words
  const uint32_t* muR;  // mu * Rn mod n
  uint32_t n_inv32;
  uint32_t n_n0inv;
  int nl;
  int xl;  // words per element of x
  uint32_t* pt;
  size_t count;
};

__global__ void raw_finish_kernel(const RawFinishParams p) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.count) return;
  uint32_t l[kMaxPrimeWords];
  exact_div_minus1(l, p.x + i * (size_t)p.xl, p.n, p.n_inv32, p.nl);
  mont_mul_serial(p.pt + i * (size_t)p.nl, l, p.muR, p.n, p.n_n0inv, p.nl);
}

// --------------------------------------------------------------------------
// Integer-pipe peak: carry-dependent IMAD.WIDE.U32 chains, the exact
// instruction the Montgomery rows are made of.
// --------------------------------------------------------------------------
constexpr int kPeakIters = 2048;
__global__ void int_peak_kernel(uint32_t* out, uint32_t a, uint32_t b) {
  uint32_t e[17], o[17];
#pragma unroll
  for (int i = 0; i < 17; i++) {
    e[i] = threadIdx.x + i;
    o[i] = threadIdx.x * 3 + i;
  }
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < kPeakIters; it++) {
    mad_lo_cc(e[0], x, y, e[0]);
    madc_hi_cc(e[1], x, y, e[1]);
#pragma unroll
    for (int i = 2; i < 16; i += 2) {
      madc_lo_cc(e[i], x, y, e[i]);
      madc_hi_cc(e[i + 1], x, y, e[i + 1]);
    }
    addc(e[16], e[16], 0);
    mad_lo_cc(o[0], y, x, o[0]);
    madc_hi_cc(o[1], y, x, o[1]);
#pragma unroll
    for (int i = 2; i < 16; i += 2) {
      madc_lo_cc(o[i], y, x, o[i]);
      madc_hi_cc(o[i + 1], y, x, o[i + 1]);
    }
    addc(o[16], o[16], 0);
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 17; i++) s ^= e[i] ^ o[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// The same peak measurement with the operand pattern of a real CIOS row: 16
// different multiplicand registers per chain (a[j] * b, n[j] * q) instead of
// one register pair reused by every multiply.
__global__ void int_peak_row_kernel(uint32_t* out, uint32_t a0, uint32_t b0) {
  uint32_t e[17], o[17], a[16], n[16];
#pragma unroll
  for (int i = 0; i < 17; i++) {
    e[i] = threadIdx.x + i;
    o[i] = threadIdx.x * 3 + i;
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    a[i] = a0 * (i + 1) + threadIdx.x;
    n[i] = b0 * (i + 3) ^ threadIdx.x;
  }
  uint32_t b = b0, q = a0 ^ 0x5555u;
  for (int it = 0; it < kPeakIters; it++) {
    mad_lo_cc(e[0], a[0], b, e[0]);
    madc_hi_cc(e[1], a[0], b, e[1]);
#pragma unroll
    for (int i = 2; i < 16; i += 2) {
      madc_lo_cc(e[i], a[i], b, e[i]);
      madc_hi_cc(e[i + 1], a[i], b, e[i + 1]);
    }
    addc(e[16], e[16], 0);
    mad_lo_cc(o[0], n[1], q, o[0]);
    madc_hi_cc(o[1], n[1], q, o[1]);
#pragma unroll
    for (int i = 2; i < 16; i += 2) {
      madc_lo_cc(o[i], n[i + 1 < 16 ? i + 1 : 15], q, o[i]);
      madc_hi_cc(o[i + 1], n[i + 1 < 16 ? i + 1 : 15], q, o[i + 1]);
    }
    addc(o[16], o[16], 0);
    b += e[16];
    q ^= o[16];
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 17; i++) s ^= e[i] ^ o[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// --------------------------------------------------------------------------
// K6: the DJN randoms of a batch drawn on the device.  The reference draws
//     r = getRandomBN(randbits) per element on the host (ipcl/pub_key.cpp:59-61,
//     ipcl/utils/common.cpp:42-101: RDSEED/RDRAND/IPP PRNG); here the host
//     supplies only a fresh 256-bit key from the OS entropy pool and the r of
//     the whole batch are the ChaCha20 keystream under it (the block function
//     of RFC 8439 section 2.3, 20 rounds): element e takes the blocks with
//     counters e*bpe .. e*bpe + bpe-1, bpe = ceil(words/16), truncated to
//     `bits` bits.  One thread per 64-byte block, 128-bit stores.
// --------------------------------------------------------------------------
struct ChachaParams {
  uint32_t key[8];
  uint32_t nonce[3];
  uint32_t* out;         // count x words
  size_t count;          // elements of this launch
  uint64_t first;        // global index of element 0 (shards of one batch)
  int words, bits, bpe;  // bpe = blocks per element
};

__device__ __forceinline__ void chacha_qr(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  a += b; d ^= a; d = __funnelshift_l(d, d, 16);
  c += d; b ^= c; b = __funnelshift_l(b, b, 12);
  a += b; d ^= a; d = __funnelshift_l(d, d, 8);
  c += d; b ^= c; b = __funnelshift_l(b, b, 7);
}

__global__ void __launch_bounds__(256) chacha20_fill_kernel(const ChachaParams p) {
  const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t e = blk / (size_t)p.bpe;
  const int j = (int)(blk % (size_t)p.bpe);
  if (e >= p.count) return;
  uint32_t in[16], x[16];
  in[0] = 0x61707865u; in[1] = 0x3320646eu; in[2] = 0x79622d32u; in[3] = 0x6b206574u;
#pragma unroll
  for (int i = 0; i < 8; i++) in[4 + i] = p.key[i];
  in[12] = (uint32_t)((p.first + e) * (uint64_t)p.bpe + (uint64_t)j);
  in[13] = p.nonce[0]; in[14] = p.nonce[1]; in[15] = p.nonce[2];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = in[i];
#pragma unroll
  for (int r = 0; r < 10; r++) {
    chacha_qr(x[0], x[4], x[8], x[12]);
    chacha_qr(x[1], x[5], x[9], x[13]);
    chacha_qr(x[2], x[6], x[10], x[14]);
    chacha_qr(x[3], x[7], x[11], x[15]);
    chacha_qr(x[0], x[5], x[10], x[15]);
    chacha_qr(x[1], x[6], x[11], x[12]);
    chacha_qr(x[2], x[7], x[8], x[13]);
    chacha_qr(x[3], x[4], x[9], x[14]);
  }
  // word w of the element is word w - 16*j of this block; bits above `bits` are zero
  uint32_t* dst = p.out + e * (size_t)p.words + (size_t)16 * j;
  const int w0 = 16 * j;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int w = w0 + i;
    uint32_t v = x[i] + in[i];
    const int left = p.bits - 32 * w;  // bits of this word that are kept
    if (left <= 0) v = 0;
    else if (left < 32) v &= (1u << left) - 1u;
    x[i] = v;
  }
  if ((p.words & 3) == 0) {
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      if (w0 + i < p.words)
        *reinterpret_cast<uint4*>(dst + i) = make_uint4(x[i], x[i + 1], x[i + 2], x[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++)
      if (w0 + i < p.words) dst[i] = x[i];
  }
}


}  // namespace ipclb200
