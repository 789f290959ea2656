// mont_core.cuh -- warp-cooperative Montgomery arithmetic for sm_100a.
//
// Replaces the arithmetic the reference gets from IPP-Crypto's mbx_exp_mb8 /
// ippsMontMul (called at ipcl/mod_exp.cpp:514 and :549-579).  One big integer
// of L = K*T 32-bit limbs is spread over a group of T adjacent lanes of a warp;
// lane t owns limbs [t*K, (t+1)*K).
//
// What was measured on B200 (profiles/r01_pipe_peaks.md) and shapes this file:
//   IMAD (32-bit)            64 lanes/clk/SM
//   IMAD.WIDE.U32(.X)        32 lanes/clk/SM   <- the only way to get a full
//                                                  32x32->64 multiply-add; ptxas
//                                                  fuses mad.lo.cc+madc.hi.cc
//   IMAD.HI                  32 lanes/clk/SM
//   SHFL                     32 lanes/clk/SM (separate pipe)
// so one MAC32 costs one IMAD.WIDE slot (4 issue cycles per warp instruction
// per SM sub-partition) and up to three other instructions ride for free in
// its shadow.  The multiply is therefore written so that EVERY product is one
// fused IMAD.WIDE on a 64-bit-aligned accumulator word, and all alignment
// work is done by register renaming:
//
//   * two accumulator arrays per lane, "even" words sitting on even limb
//     positions and "odd" words on odd limb positions;
//   * one CIOS row per limb of b:  acc += a*b_i ; q = acc0 * n0' ; acc += n*q ;
//     acc >>= 32.  The 32-bit shift swaps the roles of the two arrays; the
//     array that turns odd is shifted by one word, which costs nothing because
//     IMAD.WIDE reads its addend from word u+1 and writes word u;
//   * per row only three shuffles leave the lane: b_i broadcast, q broadcast,
//     and the one limb that crosses to the lane below;
//   * carries out of a lane's top word are kept in a private overflow word and
//     resolved once per multiply with a ballot carry look-ahead.
//
// Values are kept "almost reduced": every input and output of mont_mul is
// < R = 2^(32L) (not necessarily < n); that needs a conditional subtract only
// when the result overflows R, which is decided from one carry bit.
// tools/model_montmul.py is a bit-level Python model of exactly this scheme.
#pragma once
#include <cstdint>

namespace ipclb200 {

#define IPCLB200_FULL_MASK 0xffffffffu

// ---- carry-chain PTX wrappers (one instruction each so that no operand can
// alias an output; ptxas fuses adjacent lo/hi pairs into IMAD.WIDE.U32[.X]) ---
__device__ __forceinline__ void add_cc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void addc_cc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void addc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void sub_cc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void subc_cc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void subc(uint32_t& d, uint32_t a, uint32_t b) {
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b,
                                          uint32_t c) {
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;"
               : "=r"(d)
               : "r"(a), "r"(b), "r"(c));
}
__device__ __forceinline__ void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b,
                                           uint32_t c) {
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;"
               : "=r"(d)
               : "r"(a), "r"(b), "r"(c));
}
__device__ __forceinline__ void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b,
                                           uint32_t c) {
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;"
               : "=r"(d)
               : "r"(a), "r"(b), "r"(c));
}

// ---- the same products on 64-bit accumulator words: the halves of one 64-bit
// PTX register are a register pair from the start, which lets ptxas keep the
// loop-carried accumulators in place across a rolled loop (with 32-bit halves it
// pairs them late and re-shuffles ~50 registers per iteration) -----------------
__device__ __forceinline__ uint32_t lo32(uint64_t x) { return (uint32_t)x; }
__device__ __forceinline__ uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
__device__ __forceinline__ uint64_t pack64(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
// d = a*b + c, carry out (first of a chain)
__device__ __forceinline__ void mad_wide_cc(uint64_t& d, uint32_t a, uint32_t b, uint64_t c) {
  asm volatile(
      "{\n\t.reg .u32 cl, ch, dl, dh;\n\t"
      "mov.b64 {cl, ch}, %3;\n\t"
      "mad.lo.cc.u32 dl, %1, %2, cl;\n\t"
      "madc.hi.cc.u32 dh, %1, %2, ch;\n\t"
      "mov.b64 %0, {dl, dh};\n\t}"
      : "=l"(d)
      : "r"(a), "r"(b), "l"(c));
}
// d = a*b + c + carry in, carry out
__device__ __forceinline__ void madc_wide_cc(uint64_t& d, uint32_t a, uint32_t b, uint64_t c) {
  asm volatile(
      "{\n\t.reg .u32 cl, ch, dl, dh;\n\t"
      "mov.b64 {cl, ch}, %3;\n\t"
      "madc.lo.cc.u32 dl, %1, %2, cl;\n\t"
      "madc.hi.cc.u32 dh, %1, %2, ch;\n\t"
      "mov.b64 %0, {dl, dh};\n\t}"
      : "=l"(d)
      : "r"(a), "r"(b), "l"(c));
}
// lo half += x with carry out, hi half unchanged
__device__ __forceinline__ void add_lo_cc(uint64_t& d, uint64_t c, uint32_t x) {
  asm volatile(
      "{\n\t.reg .u32 cl, ch, dl;\n\t"
      "mov.b64 {cl, ch}, %1;\n\t"
      "add.cc.u32 dl, cl, %2;\n\t"
      "mov.b64 %0, {dl, ch};\n\t}"
      : "=l"(d)
      : "l"(c), "r"(x));
}
// {lo, hi} = {a + b, carry}
__device__ __forceinline__ void add_wide(uint64_t& d, uint32_t a, uint32_t b) {
  asm volatile(
      "{\n\t.reg .u32 dl, dh;\n\t"
      "add.cc.u32 dl, %1, %2;\n\t"
      "addc.u32 dh, 0, 0;\n\t"
      "mov.b64 %0, {dl, dh};\n\t}"
      : "=l"(d)
      : "r"(a), "r"(b));
}

// Everything below is parameterised on K (limbs per lane, even) and T (lanes
// per big integer, power of two <= 32).
template <int K, int T>
struct Mont {
  static_assert(K >= 2 && (K % 2) == 0, "K must be even");
  static_assert(T >= 1 && T <= 32 && (T & (T - 1)) == 0, "T must be 2^k");
  static constexpr int L = K * T;

  // lane position inside its group, and the group's first lane in the warp
  __device__ __forceinline__ static int lane_t() {
    return (int)(threadIdx.x & (T - 1));
  }
  __device__ __forceinline__ static int group_shift() {
    return (int)((threadIdx.x & 31) & ~(T - 1));
  }

  // ---- one CIOS row --------------------------------------------------------
  // P: array currently on even limb positions (K limbs + 1 overflow limb).
  // Q: array on odd positions whose one-word right shift is still pending:
  //    its logical word u lives in Q[2u+2..2u+3]; Q[1] (the high half of the
  //    word that fell off) belongs to limb 0, Q[K] is its overflow limb.
  // in_limb: the limb that crossed over from the lane above in the last row.
  // Returns this lane's limb that crosses to the lane below (zero in lane 0).
  __device__ __forceinline__ static uint32_t row(uint32_t (&P)[K + 1],
                                                 uint32_t (&Q)[K + 1],
                                                 const uint32_t (&a)[K],
                                                 const uint32_t (&n)[K],
                                                 uint32_t b, uint32_t in_limb,
                                                 uint32_t n0inv) {
    uint32_t t0, t1;
    add_cc(t0, Q[K], in_limb);
    addc(t1, 0, 0);
    // fold the stranded half word into limb 0; its carry enters the odd chain
    add_cc(P[0], P[0], Q[1]);
#pragma unroll
    for (int u = 0; u < K / 2 - 1; u++) {
      madc_lo_cc(Q[2 * u], a[2 * u + 1], b, Q[2 * u + 2]);
      madc_hi_cc(Q[2 * u + 1], a[2 * u + 1], b, Q[2 * u + 3]);
    }
    madc_lo_cc(Q[K - 2], a[K - 1], b, t0);
    madc_hi_cc(Q[K - 1], a[K - 1], b, t1);
    addc(Q[K], 0, 0);
    // even chain: P += a_even * b
    mad_lo_cc(P[0], a[0], b, P[0]);
    madc_hi_cc(P[1], a[0], b, P[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(P[2 * u], a[2 * u], b, P[2 * u]);
      madc_hi_cc(P[2 * u + 1], a[2 * u], b, P[2 * u + 1]);
    }
    addc(P[K], P[K], 0);
    // Montgomery quotient digit from lane 0 of the group
    uint32_t q = P[0] * n0inv;
    q = __shfl_sync(IPCLB200_FULL_MASK, q, 0, T);
    mad_lo_cc(P[0], n[0], q, P[0]);
    madc_hi_cc(P[1], n[0], q, P[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(P[2 * u], n[2 * u], q, P[2 * u]);
      madc_hi_cc(P[2 * u + 1], n[2 * u], q, P[2 * u + 1]);
    }
    addc(P[K], P[K], 0);
    mad_lo_cc(Q[0], n[1], q, Q[0]);
    madc_hi_cc(Q[1], n[1], q, Q[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(Q[2 * u], n[2 * u + 1], q, Q[2 * u]);
      madc_hi_cc(Q[2 * u + 1], n[2 * u + 1], q, Q[2 * u + 1]);
    }
    addc(Q[K], Q[K], 0);
    // P[0] is now the limb below this lane's window
    uint32_t down = __shfl_down_sync(IPCLB200_FULL_MASK, P[0], 1, T);
    return (lane_t() == T - 1) ? 0u : down;
  }

  // ---- group-wide carry resolution -----------------------------------------
  // r (K limbs per lane) has just been produced by a lane-local add whose
  // carry-out is g (0/1).  Propagates the carries across the T lanes with a
  // ballot look-ahead and returns the carry out of the top lane.
  __device__ __forceinline__ static uint32_t resolve(uint32_t (&r)[K],
                                                     uint32_t g) {
    if (T == 1) return g;
    uint32_t all = r[0];
#pragma unroll
    for (int j = 1; j < K; j++) all &= r[j];
    uint32_t bg = __ballot_sync(IPCLB200_FULL_MASK, g != 0);
    uint32_t bp = __ballot_sync(IPCLB200_FULL_MASK, all == 0xffffffffu);
    const int sh = group_shift();
    const uint32_t gm = (T == 32) ? 0xffffffffu : ((1u << T) - 1u);
    uint64_t gg = (bg >> sh) & gm;
    uint64_t pp = (bp >> sh) & gm;
    uint64_t cin = (pp + (gg << 1)) ^ pp;
    uint32_t c = (uint32_t)(cin >> lane_t()) & 1u;
    add_cc(r[0], r[0], c);
#pragma unroll
    for (int j = 1; j < K - 1; j++) addc_cc(r[j], r[j], 0);
    addc(r[K - 1], r[K - 1], 0);
    return (uint32_t)(cin >> T) & 1u;
  }

  // r += y (+1 entering lane 0 if plus_one); returns the carry out of the group.
  __device__ __forceinline__ static uint32_t group_add(uint32_t (&r)[K],
                                                       const uint32_t (&y)[K],
                                                       uint32_t plus_one) {
    uint32_t g, scratch;
    uint32_t c0 = (lane_t() == 0) ? plus_one : 0u;
    add_cc(scratch, c0, 0xffffffffu);  // CF = c0
#pragma unroll
    for (int j = 0; j < K; j++) addc_cc(r[j], r[j], y[j]);
    addc(g, 0, 0);
    return resolve(r, g);
  }

  // r = r - n if cond (per group), via r + ~n + 1 with the carry dropped.
  __device__ __forceinline__ static uint32_t cond_sub_n(uint32_t (&r)[K],
                                                        const uint32_t (&n)[K],
                                                        uint32_t cond) {
    uint32_t y[K];
#pragma unroll
    for (int j = 0; j < K; j++) y[j] = cond ? ~n[j] : 0u;
    return group_add(r, y, cond ? 1u : 0u);
  }

  // ---- Montgomery product --------------------------------------------------
  // r = a * b * R^-1 mod n (almost reduced: r < R).  a: this lane's K limbs of
  // the multiplicand; bsrc: this lane's K limbs of the multiplier (rows are fed
  // by broadcasting them lane by lane); n: this lane's limbs of the modulus.
  // All 32 lanes of the warp must call this together.
  __device__ __forceinline__ static void mul(uint32_t (&r)[K],
                                             const uint32_t (&a)[K],
                                             const uint32_t (&bsrc)[K],
                                             const uint32_t (&n)[K],
                                             uint32_t n0inv) {
    uint32_t E[K + 1], O[K + 1];
#pragma unroll
    for (int j = 0; j <= K; j++) {
      E[j] = 0;
      O[j] = 0;
    }
    uint32_t in_limb = 0;
#pragma unroll 1
    for (int s = 0; s < T; s++) {
#pragma unroll
      for (int j = 0; j < K; j += 2) {
        uint32_t b0 = __shfl_sync(IPCLB200_FULL_MASK, bsrc[j], s, T);
        uint32_t b1 = __shfl_sync(IPCLB200_FULL_MASK, bsrc[j + 1], s, T);
        in_limb = row(E, O, a, n, b0, in_limb, n0inv);
        in_limb = row(O, E, a, n, b1, in_limb, n0inv);
      }
    }
    finish(r, E, O, in_limb, n);
  }

  // Same product with the multiplier's limbs read from shared memory (bsm:
  // this group's L words, written by the caller, __syncwarp() done) instead of
  // being broadcast out of K registers: frees K registers per lane, which is
  // what lets the 32 x 2 layout keep 12 warps per SM.
  __device__ __forceinline__ static void mul_sb(uint32_t (&r)[K],
                                                const uint32_t (&a)[K],
                                                const uint32_t* bsm,
                                                const uint32_t (&n)[K],
                                                uint32_t n0inv) {
    uint32_t E[K + 1], O[K + 1];
#pragma unroll
    for (int j = 0; j <= K; j++) {
      E[j] = 0;
      O[j] = 0;
    }
    uint32_t in_limb = 0;
    // 16 rows per loop iteration (32 measured slower at K = 32: 162.7 vs
    // 159.2 ms, the body no longer sits well in the instruction cache)
    constexpr int SB_ROWS = 16;
#pragma unroll 1
    for (int i = 0; i < L; i += SB_ROWS) {
#pragma unroll
      for (int j = 0; j < SB_ROWS; j += 2) {
        const uint32_t b0 = bsm[i + j], b1 = bsm[i + j + 1];
        in_limb = row(E, O, a, n, b0, in_limb, n0inv);
        in_limb = row(O, E, a, n, b1, in_limb, n0inv);
      }
    }
    finish(r, E, O, in_limb, n);
  }
  // this lane's K limbs -> the group's shared-memory operand
  __device__ __forceinline__ static void put_sb(uint32_t* bsm, const uint32_t (&x)[K]) {
    __syncwarp();
    uint4* d = reinterpret_cast<uint4*>(bsm + lane_t() * K);
#pragma unroll
    for (int j = 0; j < K; j += 4) d[j / 4] = make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    __syncwarp();
  }

  // Assemble the two arrays into K limbs per lane, resolve the cross-lane
  // carries and bring the value back below R.
  __device__ __forceinline__ static void finish(uint32_t (&r)[K],
                                                uint32_t (&E)[K + 1],
                                                uint32_t (&O)[K + 1],
                                                uint32_t in_limb,
                                                const uint32_t (&n)[K]) {
    uint32_t t0, t1, ov;
    add_cc(t0, O[K], in_limb);
    addc(t1, 0, 0);
    add_cc(r[0], E[0], O[1]);
#pragma unroll
    for (int j = 1; j < K - 1; j++) addc_cc(r[j], E[j], O[j + 1]);
    addc_cc(r[K - 1], E[K - 1], t0);
    addc(ov, E[K], t1);
    // the overflow limb belongs to limb 0 of the lane above
    uint32_t ov_in = __shfl_up_sync(IPCLB200_FULL_MASK, ov, 1, T);
    if (lane_t() == 0) ov_in = 0;
    uint32_t g;
    add_cc(r[0], r[0], ov_in);
#pragma unroll
    for (int j = 1; j < K; j++) addc_cc(r[j], r[j], 0);
    addc(g, 0, 0);
    uint32_t top = resolve(r, g);
    // overflow past R: the top lane's own overflow limb or the resolved carry
    uint32_t bo = __ballot_sync(IPCLB200_FULL_MASK, ov != 0);
    uint32_t ovf = ((bo >> (group_shift() + T - 1)) & 1u) | top;
    if (__any_sync(IPCLB200_FULL_MASK, ovf)) cond_sub_n(r, n, ovf);
  }

  // r = (r >= n) ? r - n : r.  Returns 1 if it subtracted.
  __device__ __forceinline__ static uint32_t sub_n_if_ge(
      uint32_t (&r)[K], const uint32_t (&n)[K]) {
    uint32_t d[K], y[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
      d[j] = r[j];
      y[j] = ~n[j];
    }
    uint32_t ge = group_add(d, y, 1u);  // carry out <=> r >= n
#pragma unroll
    for (int j = 0; j < K; j++) r[j] = ge ? d[j] : r[j];
    return ge;
  }

  // Leave Montgomery form and return the canonical residue in [0, n):
  // r = x * 1 * R^-1 mod n; the product is <= n so one compare suffices.
  __device__ __forceinline__ static void from_mont(uint32_t (&r)[K],
                                                   const uint32_t (&x)[K],
                                                   const uint32_t (&n)[K],
                                                   uint32_t n0inv) {
    uint32_t one[K];
#pragma unroll
    for (int j = 0; j < K; j++) one[j] = 0;
    if (lane_t() == 0) one[0] = 1;
    mul(r, x, one, n, n0inv);
    sub_n_if_ge(r, n);
  }

  // ---- coalesced limb I/O: lane t moves its K limbs as 128-bit accesses ----
  __device__ __forceinline__ static void load(uint32_t (&x)[K],
                                              const uint32_t* __restrict__ p) {
    const uint32_t* src = p + lane_t() * K;
    if (K % 4 == 0) {
#pragma unroll
      for (int j = 0; j < K; j += 4) {
        uint4 v = *reinterpret_cast<const uint4*>(src + j);
        x[j] = v.x;
        x[j + 1] = v.y;
        x[j + 2] = v.z;
        x[j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < K; j += 2) {
        uint2 v = *reinterpret_cast<const uint2*>(src + j);
        x[j] = v.x;
        x[j + 1] = v.y;
      }
    }
  }
  __device__ __forceinline__ static void store(uint32_t* __restrict__ p,
                                               const uint32_t (&x)[K]) {
    uint32_t* dst = p + lane_t() * K;
    if (K % 4 == 0) {
#pragma unroll
      for (int j = 0; j < K; j += 4)
        *reinterpret_cast<uint4*>(dst + j) =
            make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < K; j += 2)
        *reinterpret_cast<uint2*>(dst + j) = make_uint2(x[j], x[j + 1]);
    }
  }
};

}  // namespace ipclb200
