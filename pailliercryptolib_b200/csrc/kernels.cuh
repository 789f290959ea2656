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

#include "mont_core.cuh"
#include "mont_fp64.cuh"
#include "mont_sqr.cuh"
#include "mont_tile.cuh"

namespace ipclb200 {

constexpr int kBlockThreads = 128;
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

// Dynamic work distribution: a warp claims the next chunk of 32/T consecutive
// elements from a global counter (zeroed by the host before the launch).  The
// batch is only a few waves of the resident groups, so a static split leaves
// whole SMs idle during the last wave; with claims the tail is shared.
__device__ __forceinline__ unsigned int claim_chunk(unsigned int* counter) {
  unsigned int w = 0;
  if ((threadIdx.x & 31) == 0) w = atomicAdd(counter, 1u);
  return __shfl_sync(IPCLB200_FULL_MASK, w, 0);
}

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
};

template <int K, int T>
__global__ void __launch_bounds__(kBlockThreads)
    modmul_kernel(const ModmulParams p) {
  using M = Mont<K, T>;
  constexpr int L = K * T;
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  const size_t ngroups = (size_t)gridDim.x * gpb;
  const size_t iters = (p.count + ngroups - 1) / ngroups;
  uint32_t n[K], rr[K];
  M::load(n, p.m.n);
  M::load(rr, p.m.rr);
  for (size_t it = 0; it < iters; it++) {
    const size_t inst = it * ngroups + gid;
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

// experiment: the 64-word class as 32 limbs x 2 lanes (half the shuffles and
// row bookkeeping per multiply, 246 registers -> 2 blocks per SM)
__global__ void __launch_bounds__(kBlockThreads, 2)
    decrypt_crt_k32_kernel(const DecryptCrtParams p) {
  const size_t gpb = blockDim.x / 2;
  decrypt_int_role<32, 2>(p, blockIdx.x * gpb + threadIdx.x / 2);
}

// experiment: 32 x 2 layout with the multiplier streamed from shared memory
// (Mont::mul_sb), so that the kernel fits 168 registers = 12 warps per SM
constexpr int kSbGroupWords = 68;  // 64 limbs + pad (bank spread)
template <int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB)
    decrypt_crt_k32s_kernel(const DecryptCrtParams p) {
  constexpr int K = 32, T = 2;
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  extern __shared__ uint32_t sb_smem[];
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* bsm = sb_smem + (threadIdx.x / T) * kSbGroupWords;
  uint32_t* tab = p.table_ws + gid * ((size_t)L * p.table_entries);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= 2u * nchunks) break;
    const int side = (int)(w & 1u);
    const size_t inst = (size_t)(w >> 1) * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* mn = side ? p.m1.n : p.m0.n;
    const uint32_t* mr3 = side ? p.m1.r3 : p.m0.r3;
    const uint32_t n0inv = side ? p.m1.n0inv : p.m0.n0inv;
    const uint8_t* sched = side ? p.sched1 : p.sched0;
    uint32_t n[K];
    M::load(n, mn);
    const uint32_t* c = p.ct + ii * (size_t)(2 * L);
    uint32_t acc[K];
    {
      uint32_t t[K];
      M::load(acc, c);
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = 0;
      if (M::lane_t() == 0) t[0] = 1;
      M::put_sb(bsm, t);
      M::mul_sb(acc, acc, bsm, n, n0inv);  // lo * R^-1
      M::load(t, c + L);
      uint32_t cy = M::group_add(acc, t, 0u);
      if (__any_sync(IPCLB200_FULL_MASK, cy)) M::cond_sub_n(acc, n, cy);
      M::load(t, mr3);
      M::put_sb(bsm, t);
      M::mul_sb(acc, acc, bsm, n, n0inv);  // ct * R mod n
    }
    const int nodd = sched[0];
    M::store(tab, acc);
    M::put_sb(bsm, acc);
    M::mul_sb(acc, acc, bsm, n, n0inv);  // x^2
    M::put_sb(bsm, acc);
    M::load(acc, tab);
    for (int i = 1; i < nodd; i++) {
      M::mul_sb(acc, acc, bsm, n, n0inv);
      M::store(tab + (size_t)i * L, acc);
    }
    M::load(acc, tab + (size_t)sched[1] * L);
    M::put_sb(bsm, acc);
    const uint8_t* op = sched + 2;
#pragma unroll 1
    for (uint32_t o = __ldg(op); o != 0xffu; o = __ldg(++op)) {
      // the running value is the shared-memory operand; a window multiply
      // loads the table entry into the register operand instead
      if (o) M::load(acc, tab + (size_t)(o - 1) * L);
      M::mul_sb(acc, acc, bsm, n, n0inv);
      M::put_sb(bsm, acc);
    }
    {
      uint32_t t[K];
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = 0;
      if (M::lane_t() == 0) t[0] = 1;
      M::put_sb(bsm, t);
      M::mul_sb(acc, acc, bsm, n, n0inv);
      M::sub_n_if_ge(acc, n);
    }
    if (valid) M::store(p.x + (inst * 2 + side) * L, acc);
  }
}

// --------------------------------------------------------------------------
// K4s: the CRT-decrypt modexp with symmetric squarings (mont_sqr.cuh): the
// squarings of the schedule (85 % of the products) go through MontSqr::sqr
// (1544 IMAD.WIDE instead of 2048), everything else is the integer role above.
// 64-word class only (p^2 of a 2048-bit key).
// --------------------------------------------------------------------------
__device__ __forceinline__ void modexp_sched_core_sqr(
    uint32_t (&acc)[16], const uint32_t (&xm)[16], const uint32_t (&n)[16],
    uint32_t n0inv, const uint8_t* __restrict__ sched,
    uint32_t* __restrict__ tab, uint32_t* gs, const uint32_t* zero) {
  constexpr int K = 16, T = 4;
  using M = Mont<K, T>;
  constexpr int L = K * T;
  const int nodd = sched[0];
  {
    uint32_t t[K], x2[K];
    M::store(tab, xm);
    MontSqr::sqr(x2, xm, n, n0inv, gs, zero);
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
    if (o) {
      uint32_t b[K];
      M::load(b, tab + (size_t)(o - 1) * L);
      M::mul(acc, acc, b, n, n0inv);
    } else {
      MontSqr::sqr(acc, acc, n, n0inv, gs, zero);
    }
  }
}

__global__ void __launch_bounds__(kBlockThreads, 3)
    decrypt_crt_sqr_kernel(const DecryptCrtParams p) {
  constexpr int K = 16, T = 4;
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  extern __shared__ uint32_t sqr_smem[];
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* gs = sqr_smem + (threadIdx.x / T) * kSqrGroupWords;
  uint32_t* zero = sqr_smem + gpb * kSqrGroupWords;
  MontSqr::sqr_init(sqr_smem, (int)(gpb * kSqrGroupWords + 16));
  uint32_t* tab = p.table_ws + gid * ((size_t)L * p.table_entries);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
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
      M::mul(lo, lo, t, n, n0inv);
      uint32_t cy = M::group_add(lo, hi, 0u);
      if (__any_sync(IPCLB200_FULL_MASK, cy)) M::cond_sub_n(lo, n, cy);
      M::load(t, mr3);
      M::mul(x, lo, t, n, n0inv);
    }
    modexp_sched_core_sqr(acc, x, n, n0inv, side ? p.sched1 : p.sched0, tab, gs, zero);
    M::from_mont(x, acc, n, n0inv);
    if (valid) M::store(p.x + (inst * 2 + side) * L, x);
  }
}

// the same with the 32 x 2 layout (MontSqr2): half the lanes per integer, so
// half the recombination work per integer
__global__ void __launch_bounds__(kBlockThreads, 2)
    decrypt_crt_sqr2_kernel(const DecryptCrtParams p) {
  constexpr int K = 32, T = 2;
  using M = Mont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  extern __shared__ uint32_t sqr_smem[];
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* gs = sqr_smem + (threadIdx.x / T) * kSq2GroupWords;
  uint32_t* zero = sqr_smem + gpb * kSq2GroupWords;
  MontSqr::sqr_init(sqr_smem, (int)(gpb * kSq2GroupWords + 32));
  uint32_t* tab = p.table_ws + gid * ((size_t)L * p.table_entries);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= 2u * nchunks) break;
    const int side = (int)(w & 1u);
    const size_t inst = (size_t)(w >> 1) * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* mn = side ? p.m1.n : p.m0.n;
    const uint32_t* mr3 = side ? p.m1.r3 : p.m0.r3;
    const uint32_t n0inv = side ? p.m1.n0inv : p.m0.n0inv;
    const uint8_t* sched = side ? p.sched1 : p.sched0;
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
      M::mul(lo, lo, t, n, n0inv);
      uint32_t cy = M::group_add(lo, hi, 0u);
      if (__any_sync(IPCLB200_FULL_MASK, cy)) M::cond_sub_n(lo, n, cy);
      M::load(t, mr3);
      M::mul(x, lo, t, n, n0inv);
    }
    const int nodd = sched[0];
    {
      uint32_t t[K], x2[K];
      M::store(tab, x);
      MontSqr2::sqr(x2, x, n, n0inv, gs, zero);
#pragma unroll
      for (int j = 0; j < K; j++) t[j] = x[j];
      for (int i = 1; i < nodd; i++) {
        M::mul(t, t, x2, n, n0inv);
        M::store(tab + (size_t)i * L, t);
      }
    }
    M::load(acc, tab + (size_t)sched[1] * L);
    const uint8_t* op = sched + 2;
#pragma unroll 1
    for (uint32_t o = __ldg(op); o != 0xffu; o = __ldg(++op)) {
      if (o) {
        uint32_t b[K];
        M::load(b, tab + (size_t)(o - 1) * L);
        M::mul(acc, acc, b, n, n0inv);
      } else {
        MontSqr2::sqr(acc, acc, n, n0inv, gs, zero);
      }
    }
    M::from_mont(x, acc, n, n0inv);
    if (valid) M::store(p.x + (inst * 2 + side) * L, x);
  }
}

__global__ void __launch_bounds__(kBlockThreads, 2)
    montsqr2_test_kernel(const struct MontSqrTestParams p);

// test kernel: out_sqr = MontSqr::sqr(a), out_mul = Mont::mul(a, a), one
// 64-word integer per group
struct MontSqrTestParams {
  const uint32_t* a;
  const uint32_t* n;
  uint32_t n0inv;
  uint32_t* out_sqr;
  uint32_t* out_mul;
  size_t count;
};

__global__ void __launch_bounds__(kBlockThreads, 2)
    montsqr_test_kernel(const MontSqrTestParams p) {
  constexpr int K = 16, T = 4, L = 64;
  using M = Mont<K, T>;
  extern __shared__ uint32_t sqr_smem[];
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* gs = sqr_smem + (threadIdx.x / T) * kSqrGroupWords;
  uint32_t* zero = sqr_smem + gpb * kSqrGroupWords;
  MontSqr::sqr_init(sqr_smem, (int)(gpb * kSqrGroupWords + 16));
  const bool valid = gid < p.count;
  const size_t ii = valid ? gid : p.count - 1;
  uint32_t n[K], a[K], r[K];
  M::load(n, p.n);
  M::load(a, p.a + ii * L);
  MontSqr::sqr(r, a, n, p.n0inv, gs, zero);
  if (valid) M::store(p.out_sqr + ii * L, r);
  M::mul(r, a, a, n, p.n0inv);
  if (valid) M::store(p.out_mul + ii * L, r);
}

__global__ void __launch_bounds__(kBlockThreads, 2)
    montsqr2_test_kernel(const MontSqrTestParams p) {
  constexpr int K = 32, T = 2, L = 64;
  using M = Mont<K, T>;
  extern __shared__ uint32_t sqr_smem[];
  const size_t gpb = blockDim.x / T;
  const size_t gid = blockIdx.x * gpb + threadIdx.x / T;
  uint32_t* gs = sqr_smem + (threadIdx.x / T) * kSq2GroupWords;
  uint32_t* zero = sqr_smem + gpb * kSq2GroupWords;
  MontSqr::sqr_init(sqr_smem, (int)(gpb * kSq2GroupWords + 32));
  const bool valid = gid < p.count;
  const size_t ii = valid ? gid : p.count - 1;
  uint32_t n[K], a[K], r[K];
  M::load(n, p.n);
  M::load(a, p.a + ii * L);
  MontSqr2::sqr(r, a, n, p.n0inv, gs, zero);
  if (valid) M::store(p.out_sqr + ii * L, r);
  M::mul(r, a, a, n, p.n0inv);
  if (valid) M::store(p.out_mul + ii * L, r);
}

// --------------------------------------------------------------------------
// K4f: the same CRT-decrypt modexp on the FP64 pipe (mont_fp64.cuh), and the
// dual-pipe kernel that runs both roles side by side on every SM.
//
// The FP64 role takes the same chunks from the same work counter as the
// integer role (8 ciphertexts of one side per warp: T = 4 lanes per integer in
// both) and writes the same canonical 32-bit words to p.x, so crt_finish_kernel
// does not know which pipe produced a residue.  Per chunk:
//   stage the ciphertext words in shared memory, cut them into 22-bit limbs
//   ct = lo + hi * R  ->  ct * R^-1 = mont(lo, 1) + hi  ->  * R^3  ->  ct * R
//   odd powers x, x^3, ... into this group's table (int32 limbs, L2 resident)
//   the host-built sliding-window schedule of p-1 (same bytes as the int role)
//   leave Montgomery form (result <= n), exact carry propagation by one lane,
//   repack to 32-bit words, n -> 0, store.
// --------------------------------------------------------------------------
constexpr int kFpStage = 136;  // staging words per group (128 + zero pad)

struct DecryptFpParams {
  const uint32_t* ct;  // count x 2*OW words
  FpModConst f0, f1;   // p^2, q^2
  const uint8_t *sched0, *sched1;
  uint32_t* x;  // out: count x 2 x OW words
  size_t count;
  uint32_t* table_ws;  // int32 limbs: groups x table_entries x L
  int table_entries;
  unsigned int* work_counter;  // shared with the integer role
  int debug_stage;  // 0 = off; k > 0: stop after stage k and emit the limbs
};

template <int K, int T, int OW>
__device__ __forceinline__ void decrypt_fp_role(const DecryptFpParams& p,
                                                size_t gid, double* bsm,
                                                uint32_t* stg) {
  using F = FpMont<K, T>;
  constexpr int L = K * T;
  constexpr int GW = 32 / T;
  constexpr int WPL = OW / T;  // 32-bit words of a residue per lane
  static_assert(2 * OW + 8 <= kFpStage, "staging too small");
  static_assert(L + 4 <= kFpStage, "staging too small");
  static_assert((2 * OW / T) % 4 == 0 && WPL % 4 == 0, "128-bit accesses");
  const int t = F::lane_t();
  uint32_t* tab = p.table_ws + gid * ((size_t)L * p.table_entries);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= 2u * nchunks) break;
    const int side = (int)(w & 1u);
    const size_t inst = (size_t)(w >> 1) * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const double* gn = side ? p.f1.n : p.f0.n;
    const double* gr3 = side ? p.f1.r3 : p.f0.r3;
    const uint32_t* gn32 = side ? p.f1.n32 : p.f0.n32;
    const uint32_t n0inv = side ? p.f1.n0inv : p.f0.n0inv;
    const uint8_t* sched = side ? p.sched1 : p.sched0;
    double n[K], a[K];
#pragma unroll
    for (int j = 0; j < K; j++) n[j] = __ldg(gn + t * K + j);
    // stage the 2*OW ciphertext words of this group, zero padded
    {
      const uint4* c4 =
          reinterpret_cast<const uint4*>(p.ct + ii * (size_t)(2 * OW)) +
          t * (2 * OW / T / 4);
      uint4* s4 = reinterpret_cast<uint4*>(stg) + t * (2 * OW / T / 4);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2 * OW / T / 4; j++) s4[j] = c4[j];
      if (t == 0) {
        reinterpret_cast<uint4*>(stg)[2 * OW / 4] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4*>(stg)[2 * OW / 4 + 1] = make_uint4(0, 0, 0, 0);
      }
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < K; j++)
      a[j] = fp_from_u32(fp_limb_at(stg, kFpW * (t * K + j)));
    do {
      F::put_b_one(bsm);
      F::mul(a, n, n0inv, bsm);  // lo * R^-1  (<= n)
      if (p.debug_stage == 1) break;
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int off = kFpW * (L + t * K + j);
        const uint32_t h = (off < 64 * OW) ? fp_limb_at(stg, off) : 0u;
        a[j] = __dadd_rn(a[j], fp_from_u32(h));
      }
      F::normalize(a);  // ct * R^-1 mod n, < 2n
      if (p.debug_stage == 2) break;
      F::put_b_global(bsm, gr3);
      F::mul(a, n, n0inv, bsm);  // ct * R mod n: Montgomery form
      if (p.debug_stage == 3) break;
      // odd powers
      const int nodd = sched[0];
      F::store_tab(tab, a);
      F::put_b(bsm, a);
      F::mul(a, n, n0inv, bsm);  // x^2
      if (p.debug_stage == 4) break;
      F::put_b(bsm, a);
      F::load_tab(a, tab);
      for (int i = 1; i < nodd; i++) {
        F::mul(a, n, n0inv, bsm);
        F::store_tab(tab + (size_t)i * L, a);
      }
      if (p.debug_stage == 5) break;
      F::load_tab(a, tab + (size_t)sched[1] * L);
      F::put_b(bsm, a);
      const uint8_t* op = sched + 2;
#pragma unroll 1
      for (uint32_t o = __ldg(op); o != 0xffu; o = __ldg(++op)) {
        if (o) F::load_tab(a, tab + (size_t)(o - 1) * L);
        F::mul(a, n, n0inv, bsm);
        F::put_b(bsm, a);
      }
      if (p.debug_stage == 6) break;
      F::put_b_one(bsm);
      F::mul(a, n, n0inv, bsm);  // leave Montgomery form: <= n
    } while (0);
    // limbs -> exact 22-bit digits -> 32-bit words
#pragma unroll
    for (int j = 0; j < K; j++) stg[t * K + j] = fp_to_u32(a[j]);
    if (t == 0) {
      stg[L] = 0;
      stg[L + 1] = 0;
      stg[L + 2] = 0;
    }
    __syncwarp();
    if (t == 0) {
      uint32_t c = 0;
      for (int g = 0; g < L; g++) {
        const uint32_t v = stg[g] + c;
        stg[g] = v & kFpMask;
        c = v >> kFpW;
      }
    }
    __syncwarp();
    uint32_t wv[WPL];
    bool eq = true;
#pragma unroll
    for (int j = 0; j < WPL; j++) {
      const int k = t * WPL + j;
      const int g0 = (32 * k) / kFpW;
      const int o = 32 * k - kFpW * g0;
      const uint64_t v = (uint64_t)stg[g0] | ((uint64_t)stg[g0 + 1] << kFpW) |
                         ((uint64_t)stg[g0 + 2] << (2 * kFpW));
      wv[j] = (uint32_t)(v >> o);
      eq = eq && (wv[j] == __ldg(gn32 + k));
    }
    // the product is <= n; n itself (only for a ciphertext divisible by the
    // prime) is the residue 0
    {
      const uint32_t be = __ballot_sync(IPCLB200_FULL_MASK, eq);
      const uint32_t gm = ((1u << T) - 1u) << ((threadIdx.x & 31) & ~(T - 1));
      if ((be & gm) == gm && p.debug_stage == 0) {
#pragma unroll
        for (int j = 0; j < WPL; j++) wv[j] = 0;
      }
    }
    if (valid) {
      uint4* d = reinterpret_cast<uint4*>(p.x + (inst * 2 + side) * OW) +
                 t * (WPL / 4);
#pragma unroll
      for (int j = 0; j < WPL; j += 4)
        d[j / 4] = make_uint4(wv[j], wv[j + 1], wv[j + 2], wv[j + 3]);
    }
  }
}

constexpr size_t fp_role_smem(int K, int T) {
  return (size_t)(kBlockThreads / T) * (K * T + 1) * sizeof(double) +
         (size_t)(kBlockThreads / T) * kFpStage * sizeof(uint32_t);
}

template <int K, int T, int OW>
__device__ __forceinline__ void decrypt_fp_block(const DecryptFpParams& p) {
  extern __shared__ double fp_smem[];
  constexpr int gpb = kBlockThreads / T;
  const int g = threadIdx.x / T;
  uint32_t* stg0 =
      reinterpret_cast<uint32_t*>(fp_smem + gpb * FpMont<K, T>::BSTRIDE);
  decrypt_fp_role<K, T, OW>(p, (size_t)blockIdx.x * gpb + g,
                            fp_smem + g * FpMont<K, T>::BSTRIDE,
                            stg0 + g * kFpStage);
}

// FP64 role alone: MINB = 3 -> 168 registers, MINB <= 2 -> no register cap
template <int K, int T, int OW, int MINB>
__global__ void __launch_bounds__(kBlockThreads, MINB)
    decrypt_crt_fp_kernel(const DecryptFpParams p) {
  decrypt_fp_block<K, T, OW>(p);
}

// 224 registers: one block of this kernel fits next to two blocks of the
// 144-register integer kernel on one SM (2*128*144 + 128*224 = 65536)
template <int K, int T, int OW>
__global__ void __maxnreg__(224)
    decrypt_crt_fp224_kernel(const DecryptFpParams p) {
  decrypt_fp_block<K, T, OW>(p);
}

// Both roles in one persistent kernel.  A block asks its SM for a slot number
// (per-SM atomic counter) and bit `slot` of fp_mask decides its role, so every
// SM hosts the same mix of integer-pipe and FP64-pipe warps, one warp of each
// block per SM sub-partition.
struct DecryptDualParams {
  DecryptCrtParams i;
  DecryptFpParams f;
  unsigned int* sm_slots;  // zeroed before the launch, indexed by %smid
  unsigned int fp_mask;
  unsigned int slots_per_sm;
};

template <int K, int T, int FK, int FT>
__global__ void __launch_bounds__(kBlockThreads, 3)
    decrypt_crt_dual_kernel(const DecryptDualParams p) {
  extern __shared__ double fp_smem[];
  __shared__ unsigned int s_role;
  if (threadIdx.x == 0) {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned int slot = atomicAdd(p.sm_slots + smid, 1u) % p.slots_per_sm;
    s_role = (p.fp_mask >> slot) & 1u;
  }
  __syncthreads();
  if (s_role) {
    constexpr int gpb = kBlockThreads / FT;
    const int g = threadIdx.x / FT;
    uint32_t* stg0 =
        reinterpret_cast<uint32_t*>(fp_smem + gpb * FpMont<FK, FT>::BSTRIDE);
    decrypt_fp_role<FK, FT, K * T>(p.f, (size_t)blockIdx.x * gpb + g,
                                   fp_smem + g * FpMont<FK, FT>::BSTRIDE,
                                   stg0 + g * kFpStage);
  } else {
    const size_t gpb = kBlockThreads / T;
    decrypt_int_role<K, T>(p.i, blockIdx.x * gpb + threadIdx.x / T);
  }
}

// --------------------------------------------------------------------------
// K4b: the same CRT-decrypt modexp for moduli of <= 2048 bits, one integer per
// thread (mont_tile.cuh).  A CTA works on one side (even CTAs p^2, odd CTAs
// q^2).  The whole exponentiation -- reduction of the ciphertext, table of odd
// powers, sliding-window schedule, leaving Montgomery form -- is a byte-code
// program built on the host (build_tile_program in ipcl_b200.cu), interpreted
// with ONE inlined copy of the multiply:
//   0x00        A = A^2
//   0x01..0x3f  A = A * slot[op-1]
//   0x40..0x7f  slot[op-0x40] = A
//   0x80..0xbf  A = slot[op-0x80]
//   0xc0        A = ct * R^-1      (Montgomery reduction of the 2L-limb input)
//   0xc1        A = A * R^3 / R    (-> ct * R, Montgomery form)
//   0xc2        A = A * R^-1       (leave Montgomery form), canonical, stop
// Table slots live in global memory as slot[s][v][thread] (uint4), so a warp
// reads and writes 512 contiguous bytes.
// --------------------------------------------------------------------------
struct DecryptTileParams {
  const uint32_t* ct;  // count x 2L words
  // per side (p^2, q^2): no arrays here, a runtime-indexed kernel parameter
  // would be copied to local memory
  ModConst m0, m1;
  uint4 ninv0_lo, ninv0_hi, ninv1_lo, ninv1_hi;  // -N^-1 mod 2^256
  const uint8_t *prog0, *prog1;
  uint32_t* x;  // out: count x 2 x L words
  size_t count;
  uint4* table_ws;
  int slots;
  unsigned int* work_counter;  // zeroed before the launch
};

template <int NB, int NT>
__global__ void __launch_bounds__(NT) decrypt_tile_kernel(const DecryptTileParams p) {
  using TM = TileMont<NB, NT>;
  constexpr int V = 2 * NB;
  constexpr int L = 8 * NB;
  extern __shared__ uint4 tile_smem[];
  uint4* A = tile_smem;
  uint4* QR = tile_smem + V * NT;
  uint4* s_const = tile_smem + 2 * V * NT;  // [n0 | r3_0 | n1 | r3_1], V each
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  for (int v = tid; v < V; v += NT) {
    s_const[v] = reinterpret_cast<const uint4*>(p.m0.n)[v];
    s_const[V + v] = reinterpret_cast<const uint4*>(p.m0.r3)[v];
    s_const[2 * V + v] = reinterpret_cast<const uint4*>(p.m1.n)[v];
    s_const[3 * V + v] = reinterpret_cast<const uint4*>(p.m1.r3)[v];
  }
  __syncthreads();
  // table slots of this warp: slot[s][v][lane]
  const size_t warp_global = (size_t)blockIdx.x * (NT / 32) + (tid >> 5);
  uint4* slot = p.table_ws + warp_global * ((size_t)p.slots * V * 32);
  // work items: one warp-sized chunk of one side, handed out dynamically so
  // that warps which finish early pick up the tail
  const unsigned int nchunks = (unsigned int)((p.count + 31) / 32);
  for (;;) {
    const unsigned int w = claim_chunk(p.work_counter);
    if (w >= 2u * nchunks) break;
    const int side = (int)(w & 1u);
    const size_t inst = (size_t)(w >> 1) * 32 + lane;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint4* s_n = s_const + (side ? 2 * V : 0);
    const uint4* g_r3 = reinterpret_cast<const uint4*>(side ? p.m1.r3 : p.m0.r3);
    const uint4 nl = side ? p.ninv1_lo : p.ninv0_lo, nh = side ? p.ninv1_hi : p.ninv0_hi;
    const uint32_t ninv[8] = {nl.x, nl.y, nl.z, nl.w, nh.x, nh.y, nh.z, nh.w};
    const uint4* c = reinterpret_cast<const uint4*>(p.ct + ii * (size_t)(2 * L));
    const uint8_t* pc = side ? p.prog1 : p.prog0;
#pragma unroll 1
    for (;;) {
      const uint32_t op = __ldg(pc++);
      int mode, bstride = 1;
      const uint4* Bg = nullptr;
      if (op >= 0x40u && op < 0x80u) {  // store A
        uint4* d = slot + (size_t)(op - 0x40u) * (V * 32);
        for (int v = 0; v < V; v++) d[v * 32 + lane] = A[v * NT + tid];
        continue;
      }
      if (op >= 0x80u && op < 0xc0u) {  // load A
        const uint4* d = slot + (size_t)(op - 0x80u) * (V * 32);
        for (int v = 0; v < V; v++) A[v * NT + tid] = d[v * 32 + lane];
        continue;
      }
      if (op == 0x00u) {
        mode = TM::kSqr;
      } else if (op < 0x40u) {
        Bg = slot + (size_t)(op - 1u) * (V * 32) + lane;
        bstride = 32;
        mode = TM::kMul;
      } else if (op == 0xc0u) {
        for (int v = 0; v < V; v++) A[v * NT + tid] = c[v];
        Bg = c + V;
        mode = TM::kRed;
      } else if (op == 0xc1u) {
        Bg = g_r3;
        mode = TM::kMul;
      } else {  // 0xc2
        mode = TM::kRed;
      }
      TM::mont(QR, A, Bg, bstride, s_n, ninv, mode, tid);
      uint4* t = A;
      A = QR;
      QR = t;
      if (op == 0xc2u) break;
    }
    if (TM::ge_mod(A, s_n, tid)) TM::sub_mod(A, s_n, tid);
    if (valid) {
      uint4* o = reinterpret_cast<uint4*>(p.x + (inst * 2 + side) * L);
      for (int v = 0; v < V; v++) o[v] = A[v * NT + tid];
    }
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
  // l = (dq - dp) mod q   (dp < p < q)
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
      c += (uint64_t)l[j] + p.q[j];
      l[j] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint32_t u[kMaxPrimeWords];
  mont_mul_serial(u, l, p.pinvR, p.q, p.q_n0inv, pl);
  // pt = dp + u * p   (2*pl words, < n)
  uint32_t o[kMaxPrimeWords];
  for (int j = 0; j < 2 * pl; j++) o[j] = (j < pl) ? dp[j] : 0u;
  for (int a = 0; a < pl; a++) {
    uint64_t c = 0;
    uint32_t ua = u[a];
    for (int j = 0; j < pl; j++) {
      c += (uint64_t)ua * p.p[j] + o[a + j];
      o[a + j] = (uint32_t)c;
      c >>= 32;
    }
    for (int j = a + pl; c && j < 2 * pl; j++) {
      c += o[j];
      o[j] = (uint32_t)c;
      c >>= 32;
    }
  }
  uint32_t* dst = p.pt + i * (size_t)(2 * pl);
  for (int j = 0; j < 2 * pl; j++) dst[j] = o[j];
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
// Pipe-overlap probe: do IMAD.WIDE (integer multiply pipe) and DFMA (FP64
// pipe) run at the same time on one SM sub-partition?  mode 0: every warp runs
// IMAD.WIDE carry chains; 1: every warp runs independent DFMA chains; 2: warps
// 0-3 / 8-11 / ... integer, warps 4-7 / 12-15 / ... DFMA (each sub-partition
// hosts both kinds); 3: like 2 but the DFMA warps idle (half the integer work
// alone); 4: like 2 but the integer warps idle.
// --------------------------------------------------------------------------
__global__ void pipe_mix_kernel(uint32_t* out, int mode, int iters, uint32_t a,
                                double da) {
  const int warp = threadIdx.x >> 5;
  // modes 5-8: the second role is an ALU-pipe stream instead of DFMA:
  // 5 = IMAD.WIDE warps + add-with-carry chains (IADD3.X), 6 = those chains
  // alone, 7 = IMAD.WIDE warps + independent LOP3/SHF, 8 = those alone
  if (mode >= 5) {
    const bool alu_role = (warp >> 2) & 1;
    if ((mode == 6 || mode == 8) && !alu_role) return;
    if (alu_role) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] = threadIdx.x * 7 + i;
      uint32_t x = a + threadIdx.x;
      if (mode <= 6) {
        for (int it = 0; it < iters; it++) {
          add_cc(v[0], v[0], x);
#pragma unroll
          for (int i = 1; i < 15; i++) addc_cc(v[i], v[i], x);
          addc(v[15], v[15], x);
          add_cc(v[0], v[0], v[15]);
#pragma unroll
          for (int i = 1; i < 15; i++) addc_cc(v[i], v[i], x);
          addc(v[15], v[15], x);
        }
      } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
          for (int i = 0; i < 16; i++) v[i] = __funnelshift_l(v[i], x, 3) ^ v[(i + 1) & 15];
#pragma unroll
          for (int i = 0; i < 16; i++) v[i] = (v[i] & x) | (v[(i + 5) & 15] >> 1);
        }
      }
      uint32_t s2 = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) s2 ^= v[i];
      out[blockIdx.x * blockDim.x + threadIdx.x] = s2;
      return;
    }
    mode = 3;  // integer role below, the other warps have returned
  }
  const bool fp_role = mode == 1 || (mode >= 2 && ((warp >> 2) & 1));
  if ((mode == 3 && fp_role) || (mode == 4 && !fp_role)) return;
  uint32_t s = 0;
  if (!fp_role) {
    uint32_t e[17], o[17];
#pragma unroll
    for (int i = 0; i < 17; i++) {
      e[i] = threadIdx.x + i;
      o[i] = threadIdx.x * 3 + i;
    }
    uint32_t x = a + threadIdx.x, y = 5u;
    for (int it = 0; it < iters; it++) {
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
#pragma unroll
    for (int i = 0; i < 17; i++) s ^= e[i] ^ o[i];
  } else {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = (double)(threadIdx.x + i);
    const double x = da + (double)threadIdx.x, y = 3.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = __fma_rn(x, y, acc[i]);
#pragma unroll
      for (int i = 0; i < 16; i++) acc[i] = __fma_rn(y, x, acc[i]);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) t += acc[i];
    s = (uint32_t)__double2loint(t) ^ (uint32_t)__double2hiint(t);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace ipclb200
