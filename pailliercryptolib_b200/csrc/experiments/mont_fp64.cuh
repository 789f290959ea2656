// mont_fp64.cuh -- Montgomery arithmetic on the FP64 pipe of sm_100a.
//
// Why: the integer kernels of mont_core.cuh saturate the one pipe that issues
// IMAD.WIDE (32 lanes/clk/SM, profiles/r01_pipe_peaks.md) while using a third
// of the issue slots; DFMA runs on its own pipe at 64 lanes/clk/SM and is idle.
// This file computes the same modexp (ipcl/mod_exp.cpp:514 -> mbx_exp_mb8; the
// CRT-decrypt call sites are ipcl/pri_key.cpp:128-134) with DFMA only, so that
// warps running it can share an SM with warps running the integer kernel and
// the two pipes work at the same time (decrypt_crt_dual in kernels.cuh).
//
// Representation: radix 2^22.  A big integer is L = K*T limbs, each an
// integer-valued double; lane t of a group of T lanes owns limbs [tK, (t+1)K).
// A product of two limbs is < 2^44.01 and fma() accumulates it EXACTLY as long
// as the column sum stays below 2^53; a column receives two products per row
// for at most L rows, so with L = 96 it stays below 2^51.6 and no intermediate
// normalisation is needed inside a multiply.  The row is CIOS with a one-limb
// right shift done by the fma itself (the q*n half writes column j-1):
//     acc[j]   = fma(a[j], b_i, acc[j])           j = 0..K-1
//     q        = (acc[0] mod 2^22) * n0' mod 2^22 (lane 0, integer ALU)
//     low      = fma(n[0], q, acc[0])
//     acc[j-1] = fma(n[j], q, acc[j])             j = 1..K-1
//     acc[K-1] = low of the lane above; lane 0: acc[0] += low / 2^22
// b_i comes from shared memory (one LDS.64 per row); q is broadcast as a 22-bit
// integer (one SHFL); `low` moves one lane down (two SHFL).
// Limbs are kept lazily normalised (< 2^22 + 2^9); values are almost reduced
// (<= n + 1 < 2n), R = 2^(22 L) > 2^62 n.
// tools/model_fp64_mont.py is the bit-level model of this file, bounds asserted.
#pragma once
#include <cstdint>

#include "../mont_core.cuh"

namespace ipclb200 {

constexpr int kFpW = 22;
constexpr uint32_t kFpMask = (1u << kFpW) - 1u;

// per-modulus constants of the FP64 path (device pointers)
struct FpModConst {
  const double* n;      // L limbs of the modulus
  const double* r3;     // R^3 mod n, R = 2^(22 L)
  const uint32_t* n32;  // the modulus as 32-bit words (final comparison)
  uint32_t n0inv;       // -n^-1 mod 2^22
};

__device__ __forceinline__ double fp_two52() { return 4503599627370496.0; }
// exact uint32 -> double through the 2^52 mantissa trick (one DADD)
__device__ __forceinline__ double fp_from_u32(uint32_t v) {
  return __dsub_rn(__hiloint2double(0x43300000, (int)v), fp_two52());
}
// integer-valued double in [0, 2^32) -> uint32 (one DADD)
__device__ __forceinline__ uint32_t fp_to_u32(double x) {
  return (uint32_t)__double2loint(__dadd_rn(x, fp_two52()));
}

template <int K, int T>
struct FpMont {
  static constexpr int L = K * T;
  // shared-memory operand: L doubles per group, padded so that the groups of a
  // warp read different banks when they all fetch their limb i
  static constexpr int BSTRIDE = L + 1;

  __device__ __forceinline__ static int lane_t() {
    return (int)(threadIdx.x & (T - 1));
  }

  // Lazy carry propagation, two passes: x_j <- (x_j mod 2^22) + floor(x_{j-1} /
  // 2^22).  floor to a multiple of 2^22 = add 2^74 rounding down, subtract it.
  // In: limbs < 2^52.  Out: limbs < 2^22 + 2^9.
  __device__ __forceinline__ static void normalize(double (&x)[K]) {
    const double MW = 18889465931478580854784.0;   // 2^74
    const double INVW = 2.384185791015625e-07;     // 2^-22
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
      double h_prev = 0.0;
#pragma unroll
      for (int j = 0; j < K; j++) {
        const double h = __dsub_rn(__dadd_rd(x[j], MW), MW);
        const double lo = __dsub_rn(x[j], h);
        x[j] = __fma_rn(h_prev, INVW, lo);
        h_prev = h;
      }
      double up = __shfl_up_sync(IPCLB200_FULL_MASK, h_prev, 1, T);
      if (lane_t() == 0) up = 0.0;
      x[0] = __fma_rn(up, INVW, x[0]);
    }
  }

  // a <- a * b * R^-1 mod n (almost reduced, lazily normalised).  b: the L
  // limbs of the multiplier in shared memory (this group's buffer).  All 32
  // lanes of the warp must call this together.
  __device__ __forceinline__ static void mul(double (&a)[K],
                                             const double (&n)[K],
                                             uint32_t n0inv,
                                             const double* bsm) {
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; j++) acc[j] = 0.0;
    const double cmul = (lane_t() == 0) ? 2.384185791015625e-07 : 0.0;
    const bool top = lane_t() == T - 1;
#pragma unroll 2
    for (int i = 0; i < L; i++) {
      const double b = bsm[i];
      acc[0] = __fma_rn(a[0], b, acc[0]);
      // Montgomery quotient digit from lane 0 (acc[0] < 2^52: the low mantissa
      // word of acc[0] + 2^52 is acc[0] mod 2^32)
      uint32_t qi = ((uint32_t)__double2loint(__dadd_rn(acc[0], fp_two52())) *
                     n0inv) & kFpMask;
      qi = __shfl_sync(IPCLB200_FULL_MASK, qi, 0, T);
      const double q = fp_from_u32(qi);
#pragma unroll
      for (int j = 1; j < K; j++) acc[j] = __fma_rn(a[j], b, acc[j]);
      const double low = __fma_rn(n[0], q, acc[0]);
#pragma unroll
      for (int j = 1; j < K; j++) acc[j - 1] = __fma_rn(n[j], q, acc[j]);
      const double in = __shfl_down_sync(IPCLB200_FULL_MASK, low, 1, T);
      acc[K - 1] = top ? 0.0 : in;
      acc[0] = __fma_rn(low, cmul, acc[0]);
    }
    normalize(acc);
#pragma unroll
    for (int j = 0; j < K; j++) a[j] = acc[j];
  }

  // this lane's K limbs <-> the group's shared-memory operand buffer
  __device__ __forceinline__ static void put_b(double* bsm,
                                               const double (&x)[K]) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; j++) bsm[lane_t() * K + j] = x[j];
    __syncwarp();
  }
  __device__ __forceinline__ static void put_b_one(double* bsm) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; j++)
      bsm[lane_t() * K + j] = (lane_t() == 0 && j == 0) ? 1.0 : 0.0;
    __syncwarp();
  }
  __device__ __forceinline__ static void put_b_global(
      double* bsm, const double* __restrict__ g) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; j++)
      bsm[lane_t() * K + j] = __ldg(g + lane_t() * K + j);
    __syncwarp();
  }

  // window-table entries live in global memory (L2) as int32 limbs, L per
  // entry; a lane only ever reads back the limbs it wrote itself
  __device__ __forceinline__ static void store_tab(uint32_t* __restrict__ e,
                                                   const double (&x)[K]) {
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    uint4* d = reinterpret_cast<uint4*>(e + lane_t() * K);
#pragma unroll
    for (int j = 0; j < K; j += 4)
      d[j / 4] = make_uint4(fp_to_u32(x[j]), fp_to_u32(x[j + 1]),
                            fp_to_u32(x[j + 2]), fp_to_u32(x[j + 3]));
  }
  __device__ __forceinline__ static void load_tab(double (&x)[K],
                                                  const uint32_t* e) {
    const uint4* s = reinterpret_cast<const uint4*>(e + lane_t() * K);
#pragma unroll
    for (int j = 0; j < K; j += 4) {
      const uint4 v = s[j / 4];
      x[j] = fp_from_u32(v.x);
      x[j + 1] = fp_from_u32(v.y);
      x[j + 2] = fp_from_u32(v.z);
      x[j + 3] = fp_from_u32(v.w);
    }
  }
};

// 22-bit limb starting at bit `off` of a little-endian word array in shared
// memory (the array is zero padded past its last word)
__device__ __forceinline__ uint32_t fp_limb_at(const uint32_t* w, int off) {
  const int idx = off >> 5, sh = off & 31;
  const uint64_t v = ((uint64_t)w[idx + 1] << 32) | w[idx];
  return (uint32_t)(v >> sh) & kFpMask;
}

}  // namespace ipclb200
