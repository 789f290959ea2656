// mont_tile.cuh -- thread-per-integer Montgomery arithmetic for moduli of up
// to 2048 bits (the p^2 / q^2 side of CRT decryption at keys <= 2048 bits).
//
// Why a second formulation next to mont_core.cuh: in a modexp ~85 % of the
// products are squarings, and a square has only half as many distinct limb
// products.  The lane-distributed CIOS of mont_core.cuh cannot use that (the
// redundant products sit in different lanes; exchanging them costs more
// shuffles than it saves multiplies).  With one integer per thread nothing has
// to be exchanged: the operands live in shared memory as per-thread columns,
// the accumulators in registers, and
//
//   square  = 28 off-diagonal 8x8-limb tiles (doubled) + 8 diagonal tiles
//           + 64 reduction tiles + 8 quotient blocks          = 6688 MAC32
//   product = 64 + 64 tiles + 8 quotient blocks                = 8480 MAC32
//
// against 8192 for either in the CIOS form: ~16 % fewer IMAD.WIDE over a whole
// exponentiation, no shuffles, no ballots.
//
// Algorithm: block-level finely integrated product scanning (FIPS) with
// 8-limb blocks.  For every column block c of the 2L-limb product all 8x8 tiles
// (I, J), I + J = c, of the operand product and of Q*N are accumulated
// column-wise into fifteen 96-bit column accumulators (IMAD.WIDE.U32 + one
// IADD3.X per multiply, every column an independent chain); the quotient block
// is Q_c = low8(W * N') with N' = -N^-1 mod 2^256; the low eight limbs are then
// resolved, stored, and the window moves up by one block.  Results are almost
// reduced (< R) exactly as in mont_core.cuh.  tools/model_tile_fips.py is the
// bit-level Python model.
#pragma once
#include <cstdint>

#include "mont_core.cuh"

namespace ipclb200 {

// fifteen 96-bit column accumulators.  The low 64 bits are ONE 64-bit
// register so that ptxas keeps them in an aligned pair (separate lo[]/hi[]
// arrays made it shuffle registers with IMAD.MOV on the multiplier pipe).
struct ColAcc {
  uint64_t lh[15];
  uint32_t ex[15];
};

__device__ __forceinline__ void colacc_zero(ColAcc& w) {
#pragma unroll
  for (int k = 0; k < 15; k++) {
    w.lh[k] = 0;
    w.ex[k] = 0;
  }
}

// (ex:lh) += x * y : IMAD.WIDE.U32 with carry-out plus one IADD3.X
__device__ __forceinline__ void mac96(uint64_t& lh, uint32_t& ex, uint32_t x,
                                      uint32_t y) {
  asm volatile(
      "{\n\t"
      ".reg .u32 lo, hi;\n\t"
      "mov.b64 {lo, hi}, %0;\n\t"
      "mad.lo.cc.u32 lo, %2, %3, lo;\n\t"
      "madc.hi.cc.u32 hi, %2, %3, hi;\n\t"
      "addc.u32 %1, %1, 0;\n\t"
      "mov.b64 %0, {lo, hi};\n\t"
      "}"
      : "+l"(lh), "+r"(ex)
      : "r"(x), "r"(y));
}

// (ex:lh) += (b_ex:b_lh)
__device__ __forceinline__ void add96(uint64_t& lh, uint32_t& ex, uint64_t b_lh,
                                      uint32_t b_ex) {
  asm volatile(
      "add.cc.u64 %0, %0, %2;\n\t"
      "addc.u32 %1, %1, %3;"
      : "+l"(lh), "+r"(ex)
      : "l"(b_lh), "r"(b_ex));
}

// w += X * Y (8 x 8 limbs), column k = sum of x_i*y_j with i + j = k
__device__ __forceinline__ void tile_mac(ColAcc& w, const uint32_t (&X)[8],
                                         const uint32_t (&Y)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) mac96(w.lh[i + j], w.ex[i + j], X[i], Y[j]);
  }
}

// w += 2 * s
__device__ __forceinline__ void colacc_add_doubled(ColAcc& w, const ColAcc& s) {
#pragma unroll
  for (int k = 0; k < 15; k++)
    add96(w.lh[k], w.ex[k], s.lh[k] << 1,
          (s.ex[k] << 1) | (uint32_t)(s.lh[k] >> 63));
}

// exact low eight limbs of the window; their carry moves into column 8 and
// columns 0..7 are cleared
__device__ __forceinline__ void colacc_resolve_low(ColAcc& w, uint32_t (&out)[8]) {
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint64_t t = w.lh[k];
    uint32_t e = w.ex[k];
    add96(t, e, c, 0);
    out[k] = (uint32_t)t;
    c = (t >> 32) | ((uint64_t)e << 32);
    w.lh[k] = 0;
    w.ex[k] = 0;
  }
  add96(w.lh[8], w.ex[8], c, 0);
}

// window moves up by one block
__device__ __forceinline__ void colacc_shift(ColAcc& w) {
#pragma unroll
  for (int k = 0; k < 7; k++) {
    w.lh[k] = w.lh[k + 8];
    w.ex[k] = w.ex[k + 8];
  }
#pragma unroll
  for (int k = 7; k < 15; k++) {
    w.lh[k] = 0;
    w.ex[k] = 0;
  }
}

// q = low eight limbs of t * ninv  (36 multiplies)
__device__ __forceinline__ void low_mul8(uint32_t (&q)[8], const uint32_t (&t)[8],
                                         const uint32_t (&ninv)[8]) {
  uint64_t a = 0;
  uint32_t e = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
#pragma unroll
    for (int i = 0; i <= k; i++) mac96(a, e, t[i], ninv[k - i]);
    q[k] = (uint32_t)a;
    a = (a >> 32) | ((uint64_t)e << 32);
    e = 0;
  }
}

// Per-thread big integers in shared memory: limb vector v (four limbs) of
// thread t sits at uint4 index v*NT + t, so a warp's 128-bit access is
// conflict free.  NB = number of 8-limb blocks, NT = threads per CTA.
template <int NB, int NT>
struct TileMont {
  static constexpr int L = NB * 8;
  static constexpr int V = NB * 2;  // uint4 vectors per integer

  // shared-memory accesses through 32-bit shared-window addresses (one IMAD
  // per block instead of 64-bit generic pointer arithmetic)
  __device__ __forceinline__ static uint32_t saddr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  // block `blk` of an integer whose vectors are `stride` uint4 apart
  __device__ __forceinline__ static void ld_block(uint32_t (&x)[8], uint32_t base,
                                                  int blk, int stride) {
    uint32_t a = base + (uint32_t)(2 * blk * stride) * 16u;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]) : "r"(a));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
                 : "r"(a + (uint32_t)stride * 16u));
  }
  __device__ __forceinline__ static void st_block(uint32_t base, int blk, int stride,
                                                  const uint32_t (&x)[8]) {
    uint32_t a = base + (uint32_t)(2 * blk * stride) * 16u;
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "r"(a), "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "r"(a + (uint32_t)stride * 16u), "r"(x[4]), "r"(x[5]), "r"(x[6]),
                    "r"(x[7]) : "memory");
  }

  // w += sum over I in [i0, i1) of X_I * Y_(c-I); operands of the next tile are
  // fetched while the current one is multiplied
  __device__ __forceinline__ static void tiles(ColAcc& w, uint32_t xb, int xs,
                                               uint32_t yb, int ys, int c, int i0,
                                               int i1) {
    if (i0 >= i1) return;
    uint32_t X[8], Y[8];
    ld_block(X, xb, i0, xs);
    ld_block(Y, yb, c - i0, ys);
#pragma unroll 1
    for (int I = i0; I < i1; I++) {
      uint32_t Xn[8], Yn[8];
      const int In = I + 1 < i1 ? I + 1 : I;
      ld_block(Xn, xb, In, xs);
      ld_block(Yn, yb, c - In, ys);
      tile_mac(w, X, Y);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        X[k] = Xn[k];
        Y[k] = Yn[k];
      }
    }
  }

  enum Mode { kSqr = 0, kMul = 1, kRed = 2 };

  // QR = A*B/R (kMul), A*A/R (kSqr) or (A + B*R)/R (kRed: A, B are the low and
  // high halves of a 2L-limb number).  A, B, QR: this thread's columns in
  // shared memory (QR holds the quotient blocks first and the result after).
  // nmod: the modulus, CTA-shared plain limbs; ninv: -N^-1 mod 2^256.
  __device__ __forceinline__ static void mont(uint4* QRp, const uint4* Ap,
                                              const uint4* Bp, const uint4* nmodp,
                                              const uint32_t (&ninv)[8], int mode,
                                              int tid) {
    const uint32_t A = saddr(Ap + tid), B = saddr(Bp + tid), QR = saddr(QRp + tid);
    const uint32_t nmod = saddr(nmodp);
    ColAcc W;
    colacc_zero(W);
#pragma unroll 1
    for (int c = 0; c < 2 * NB; c++) {
      if (mode == kSqr) {
        ColAcc S;
        colacc_zero(S);
        tiles(S, A, NT, A, NT, c, c - NB + 1 > 0 ? c - NB + 1 : 0, (c + 1) >> 1);
        colacc_add_doubled(W, S);
      }
      {
        // tiles accumulated straight into W: the diagonal tile of a square,
        // every operand tile of a product, nothing for a pure reduction
        int i0, i1;
        uint32_t yarr = B;
        if (mode == kSqr) {
          i0 = c >> 1;
          i1 = ((c & 1) == 0 && (c >> 1) < NB) ? i0 + 1 : i0;
          yarr = A;
        } else if (mode == kMul) {
          i0 = c - NB + 1 > 0 ? c - NB + 1 : 0;
          i1 = (c < NB - 1 ? c : NB - 1) + 1;
        } else {
          i0 = i1 = 0;
          uint32_t X[8];
          ld_block(X, c < NB ? A : B, c < NB ? c : c - NB, NT);
#pragma unroll
          for (int k = 0; k < 8; k++) add96(W.lh[k], W.ex[k], (uint64_t)X[k], 0);
        }
        tiles(W, A, NT, yarr, NT, c, i0, i1);
      }
      // Q*N tiles of this column block except the one that needs Q_c itself
      tiles(W, QR, NT, nmod, 1, c, c < NB ? 0 : c - NB + 1, c < NB ? c : NB);
      uint32_t T[8];
      colacc_resolve_low(W, T);
      if (c < NB) {
        uint32_t Qc[8], Y[8];
        low_mul8(Qc, T, ninv);
        st_block(QR, c, NT, Qc);
#pragma unroll
        for (int k = 0; k < 8; k++) W.lh[k] = T[k];
        ld_block(Y, nmod, 0, 1);
        tile_mac(W, Qc, Y);
        colacc_resolve_low(W, T);  // all zero by construction
      } else {
        st_block(QR, c - NB, NT, T);
      }
      colacc_shift(W);
    }
    // value was < R + N: bring it back below R
    if (W.lh[0] != 0) sub_mod(QRp, nmodp, tid);
  }

  // x -= n (mod 2^(32L))
  __device__ __forceinline__ static void sub_mod(uint4* x, const uint4* nmod, int tid) {
    uint32_t borrow = 0;
#pragma unroll 1
    for (int v = 0; v < V; v++) {
      uint4 a = x[v * NT + tid];
      uint4 m = nmod[v];
      uint32_t r0, r1, r2, r3;
      sub_cc(r0, 0, borrow);  // re-seed the borrow flag: 0 - borrow underflows iff borrow
      subc_cc(r0, a.x, m.x);
      subc_cc(r1, a.y, m.y);
      subc_cc(r2, a.z, m.z);
      subc_cc(r3, a.w, m.w);
      uint32_t nb;
      subc(nb, 0, 0);  // 0 - 0 - borrow -> 0 or 0xffffffff
      borrow = nb & 1u;
      x[v * NT + tid] = make_uint4(r0, r1, r2, r3);
    }
  }

  // x >= n ?  (x, n: L limbs)
  __device__ __forceinline__ static bool ge_mod(const uint4* x, const uint4* nmod, int tid) {
    bool ge = true;  // equal counts as >=
#pragma unroll 1
    for (int v = V - 1; v >= 0; v--) {
      uint4 a = x[v * NT + tid];
      uint4 m = nmod[v];
      if (a.w != m.w) return a.w > m.w;
      if (a.z != m.z) return a.z > m.z;
      if (a.y != m.y) return a.y > m.y;
      if (a.x != m.x) return a.x > m.x;
    }
    return ge;
  }
};

}  // namespace ipclb200
