// kernels_exp.cuh -- experiments that measured slower than the shipped kernels
// (DESIGN.md section 3.6): FP64-pipe decrypt, dual-pipe decrypt, symmetric
// squarings, 32 x 2 layouts, thread-per-integer decrypt, the pipe-overlap probe.
// Compiled only with -DIPCLB200_EXPERIMENTS (build.py --experiments); the
// default libipcl_b200.so holds none of this.
#pragma once
#include "../kernels.cuh"
#include "mont_fp64.cuh"
#include "mont_sqr.cuh"
#include "mont_tile.cuh"

namespace ipclb200 {
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
