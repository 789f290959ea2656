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
// against 8192 for either in the CIOS form: ~15 % fewer IMAD.WIDE over a whole
// exponentiation, no shuffles, no ballots.
//
// Algorithm: block-level finely integrated product scanning (FIPS) with
// 8-limb blocks.  For every column block c of the 2L-limb product all 8x8 tiles
// (I, J), I + J = c, of the operand product and of Q*N are accumulated into a
// 16-limb window; the quotient block is Q_c = low8(W * N') with
// N' = -N^-1 mod 2^256; the low eight limbs are then resolved, stored, and the
// window moves up by one block.  Results are almost reduced (< R) exactly as in
// mont_core.cuh.
//
// The window (v2; v1 used 96-bit column accumulators, one IADD3.X per multiply,
// and was issue-bound at 2.7 instructions per multiply): even 64-bit words at
// even limb positions and odd words at odd positions, like mont_core.cuh; a
// tile row x_i * Y is two carry chains of four IMAD.WIDE.U32.X (even and odd
// limbs of Y), whose carry-outs go to 32-bit counters: 64 IMAD.WIDE + 16 adds
// per tile.  tools/model_tile_fips.py models the block algorithm,
// tools/model_tile_eo.py the window.
//
// Status (round 1, B200, 2048-bit key, 65536 ciphertexts): bit-exact, 165 ms
// against 159.5 ms for the lane-distributed kernel -- 15 % fewer multiplies but
// only 72 % IMAD.WIDE issue utilisation (12 warps/SM fit the shared-memory
// columns; ptxas spends another 10 % of the multiplier pipe on IMAD.MOV/IMAD.X
// register traffic).  Opt-in with IPCLB200_DECRYPT=tile; see DESIGN.md 3.6.
#pragma once
#include <cstdint>

#include "../mont_core.cuh"

namespace ipclb200 {

// value = sum E[u] 2^(64u) + O[u] 2^(64u+32) + cE[u] 2^(64u) + cO[u] 2^(64u+32)
struct Win {
  uint64_t E[8], O[8];
  uint32_t cE[9], cO[9];
};

__device__ __forceinline__ void win_zero(Win& w) {
#pragma unroll
  for (int u = 0; u < 8; u++) {
    w.E[u] = 0;
    w.O[u] = 0;
  }
#pragma unroll
  for (int u = 0; u < 9; u++) {
    w.cE[u] = 0;
    w.cO[u] = 0;
  }
}

// (w3:w2:w1:w0) += x * (y3, y2, y1, y0) as one carry chain; carry-out -> cnt
__device__ __forceinline__ void chain4(uint64_t& w0, uint64_t& w1, uint64_t& w2,
                                       uint64_t& w3, uint32_t& cnt, uint32_t x,
                                       uint32_t y0, uint32_t y1, uint32_t y2,
                                       uint32_t y3) {
  asm volatile(
      "{\n\t"
      ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
      "mov.b64 {l0, h0}, %0;\n\t"
      "mov.b64 {l1, h1}, %1;\n\t"
      "mov.b64 {l2, h2}, %2;\n\t"
      "mov.b64 {l3, h3}, %3;\n\t"
      "mad.lo.cc.u32 l0, %5, %6, l0;\n\t"
      "madc.hi.cc.u32 h0, %5, %6, h0;\n\t"
      "madc.lo.cc.u32 l1, %5, %7, l1;\n\t"
      "madc.hi.cc.u32 h1, %5, %7, h1;\n\t"
      "madc.lo.cc.u32 l2, %5, %8, l2;\n\t"
      "madc.hi.cc.u32 h2, %5, %8, h2;\n\t"
      "madc.lo.cc.u32 l3, %5, %9, l3;\n\t"
      "madc.hi.cc.u32 h3, %5, %9, h3;\n\t"
      "addc.u32 %4, %4, 0;\n\t"
      "mov.b64 %0, {l0, h0};\n\t"
      "mov.b64 %1, {l1, h1};\n\t"
      "mov.b64 %2, {l2, h2};\n\t"
      "mov.b64 %3, {l3, h3};\n\t"
      "}"
      : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+r"(cnt)
      : "r"(x), "r"(y0), "r"(y1), "r"(y2), "r"(y3));
}

// w += X * Y (8 x 8 limbs)
__device__ __forceinline__ void tile_mac(Win& w, const uint32_t (&X)[8],
                                         const uint32_t (&Y)[8]) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((i & 1) == 0) {
      const int b = i / 2;
      chain4(w.E[b], w.E[b + 1], w.E[b + 2], w.E[b + 3], w.cE[b + 4], X[i], Y[0],
             Y[2], Y[4], Y[6]);
      chain4(w.O[b], w.O[b + 1], w.O[b + 2], w.O[b + 3], w.cO[b + 4], X[i], Y[1],
             Y[3], Y[5], Y[7]);
    } else {
      const int b = i / 2, b1 = (i + 1) / 2;
      chain4(w.O[b], w.O[b + 1], w.O[b + 2], w.O[b + 3], w.cO[b + 4], X[i], Y[0],
             Y[2], Y[4], Y[6]);
      chain4(w.E[b1], w.E[b1 + 1], w.E[b1 + 2], w.E[b1 + 3], w.cE[b1 + 4], X[i],
             Y[1], Y[3], Y[5], Y[7]);
    }
  }
}

// w += P (eight limbs at limb 0)
__device__ __forceinline__ void win_add_block(Win& w, const uint32_t (&P)[8]) {
  uint64_t p0 = P[0] | ((uint64_t)P[1] << 32), p1 = P[2] | ((uint64_t)P[3] << 32);
  uint64_t p2 = P[4] | ((uint64_t)P[5] << 32), p3 = P[6] | ((uint64_t)P[7] << 32);
  asm volatile(
      "add.cc.u64 %0, %0, %5;\n\t"
      "addc.cc.u64 %1, %1, %6;\n\t"
      "addc.cc.u64 %2, %2, %7;\n\t"
      "addc.cc.u64 %3, %3, %8;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+l"(w.E[0]), "+l"(w.E[1]), "+l"(w.E[2]), "+l"(w.E[3]), "+r"(w.cE[4])
      : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}

// w += P * 2^256 (eight limbs at limb 8)
__device__ __forceinline__ void win_add_block_hi(Win& w, const uint32_t (&P)[8]) {
  uint64_t p0 = P[0] | ((uint64_t)P[1] << 32), p1 = P[2] | ((uint64_t)P[3] << 32);
  uint64_t p2 = P[4] | ((uint64_t)P[5] << 32), p3 = P[6] | ((uint64_t)P[7] << 32);
  asm volatile(
      "add.cc.u64 %0, %0, %5;\n\t"
      "addc.cc.u64 %1, %1, %6;\n\t"
      "addc.cc.u64 %2, %2, %7;\n\t"
      "addc.cc.u64 %3, %3, %8;\n\t"
      "addc.u32 %4, %4, 0;"
      : "+l"(w.E[4]), "+l"(w.E[5]), "+l"(w.E[6]), "+l"(w.E[7]), "+r"(w.cE[8])
      : "l"(p0), "l"(p1), "l"(p2), "l"(p3));
}

// exact low eight limbs of the window; they are cleared and their carry moves
// to limb 8
__device__ __forceinline__ void win_resolve_low(Win& w, uint32_t (&T)[8]) {
  uint64_t carry = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint64_t s = carry;
    if ((k & 1) == 0) {
      s += (uint32_t)w.E[k / 2];
      s += w.cE[k / 2];
      if (k) s += (uint32_t)(w.O[k / 2 - 1] >> 32);
    } else {
      s += (uint32_t)(w.E[k / 2] >> 32);
      s += (uint32_t)w.O[k / 2];
      s += w.cO[k / 2];
    }
    T[k] = (uint32_t)s;
    carry = s >> 32;
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    w.E[u] = 0;
    w.cE[u] = 0;
    w.cO[u] = 0;
  }
#pragma unroll
  for (int u = 0; u < 3; u++) w.O[u] = 0;
  w.O[3] &= 0xffffffff00000000ull;  // its upper half is limb 8
  uint64_t t = (uint64_t)w.cE[4] + carry;
  w.cE[4] = (uint32_t)t;
  w.cO[4] += (uint32_t)(t >> 32);
}

__device__ __forceinline__ void win_set_low(Win& w, const uint32_t (&T)[8]) {
#pragma unroll
  for (int u = 0; u < 4; u++) w.E[u] = T[2 * u] | ((uint64_t)T[2 * u + 1] << 32);
}

// window moves up by one block (the low eight limbs must be resolved)
__device__ __forceinline__ void win_shift(Win& w) {
  const uint32_t x = (uint32_t)(w.O[3] >> 32);
#pragma unroll
  for (int u = 0; u < 4; u++) {
    w.E[u] = w.E[u + 4];
    w.O[u] = w.O[u + 4];
    w.E[u + 4] = 0;
    w.O[u + 4] = 0;
  }
#pragma unroll
  for (int u = 0; u < 5; u++) {
    w.cE[u] = w.cE[u + 4];
    w.cO[u] = w.cO[u + 4];
  }
#pragma unroll
  for (int u = 5; u < 9; u++) {
    w.cE[u] = 0;
    w.cO[u] = 0;
  }
  uint64_t t = (uint64_t)w.cE[0] + x;
  w.cE[0] = (uint32_t)t;
  w.cO[0] += (uint32_t)(t >> 32);
}

// (ex:lh) += x * y
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

// q = low eight limbs of t * ninv  (36 multiplies, column-wise)
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

  __device__ __forceinline__ static uint32_t saddr(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  // block `blk` of a shared-memory integer whose vectors are `stride` uint4 apart
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
  // same from global memory (table slots, R^3, the high half of a ciphertext)
  __device__ __forceinline__ static void ld_block_g(uint32_t (&x)[8], const uint4* base,
                                                    int blk, int stride) {
    uint4 a = base[(2 * blk) * stride];
    uint4 b = base[(2 * blk + 1) * stride];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w;
    x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  }

  // w += sum over I in [i0, i1) of X_I * Y_(c-I).  X in shared memory; Y in
  // shared memory (yg == nullptr) or global memory.  (Fetching the next
  // tile's operands early was tried: the register copies it needs cost more
  // than the latency it hides with three warps per scheduler.)
  __device__ __forceinline__ static void tiles(Win& w, uint32_t xb, int xs, uint32_t yb,
                                               const uint4* yg, int ys, int c, int i0,
                                               int i1) {
#pragma unroll 1
    for (int I = i0; I < i1; I++) {
      uint32_t X[8], Y[8];
      ld_block(X, xb, I, xs);
      if (yg) ld_block_g(Y, yg, c - I, ys); else ld_block(Y, yb, c - I, ys);
      tile_mac(w, X, Y);
    }
  }

  // w += 2 * sum over I in [i0, i1) of A_I * A_(c-I): the off-diagonal tiles of
  // a square.  The factor two is applied to the eight-limb block A_I itself:
  // 2*A_I = (A_I << 1 mod 2^256) + t * 2^256 with t its top bit, so the tile is
  // multiplied with the shifted block and A_(c-I) is added one block higher
  // when t is set -- no second window to zero, double and add per column block.
  __device__ __forceinline__ static void tiles_dbl(Win& w, uint32_t ab, int c, int i0,
                                                   int i1) {
#pragma unroll 1
    for (int I = i0; I < i1; I++) {
      uint32_t X[8], Y[8], D[8];
      ld_block(X, ab, I, NT);
      ld_block(Y, ab, c - I, NT);
      D[0] = X[0] << 1;
#pragma unroll
      for (int k = 1; k < 8; k++) D[k] = __funnelshift_l(X[k - 1], X[k], 1);
      tile_mac(w, D, Y);
      const uint32_t mask = 0u - (X[7] >> 31);
#pragma unroll
      for (int k = 0; k < 8; k++) X[k] = Y[k] & mask;
      win_add_block_hi(w, X);
    }
  }

  enum Mode { kSqr = 0, kMul = 1, kRed = 2 };

  // QR = A*B/R (kMul), A*A/R (kSqr) or (A + B*R)/R (kRed: A and B are the low
  // and high halves of a 2L-limb number; Bg == nullptr means B = 0).
  // A, QR: this thread's columns in shared memory (QR holds the quotient blocks
  // first and the result after).  Bg: operand in global memory, vectors
  // `bstride` uint4 apart.  nmod: the modulus, CTA-shared plain limbs in shared
  // memory; ninv: -N^-1 mod 2^256.
  __device__ __forceinline__ static void mont(uint4* QRp, const uint4* Ap,
                                              const uint4* Bg, int bstride,
                                              const uint4* nmodp,
                                              const uint32_t (&ninv)[8], int mode,
                                              int tid) {
    const uint32_t A = saddr(Ap + tid), QR = saddr(QRp + tid);
    const uint32_t nmod = saddr(nmodp);
    Win W;
    win_zero(W);
#pragma unroll 1
    for (int c = 0; c < 2 * NB; c++) {
      if (mode == kSqr) {
        tiles_dbl(W, A, c, c - NB + 1 > 0 ? c - NB + 1 : 0, (c + 1) >> 1);
        // diagonal tile
        const int d = c >> 1;
        tiles(W, A, NT, A, nullptr, NT, c, d, ((c & 1) == 0 && d < NB) ? d + 1 : d);
      } else if (mode == kMul) {
        tiles(W, A, NT, 0, Bg, bstride, c, c - NB + 1 > 0 ? c - NB + 1 : 0,
              (c < NB - 1 ? c : NB - 1) + 1);
      } else {
        uint32_t X[8];
        if (c < NB) {
          ld_block(X, A, c, NT);
          win_add_block(W, X);
        } else if (Bg) {
          ld_block_g(X, Bg, c - NB, bstride);
          win_add_block(W, X);
        }
      }
      // Q*N tiles of this column block except the one that needs Q_c itself
      tiles(W, QR, NT, nmod, nullptr, 1, c, c < NB ? 0 : c - NB + 1, c < NB ? c : NB);
      uint32_t T[8];
      win_resolve_low(W, T);
      if (c < NB) {
        uint32_t Qc[8], Y[8];
        low_mul8(Qc, T, ninv);
        st_block(QR, c, NT, Qc);
        win_set_low(W, T);
        ld_block(Y, nmod, 0, 1);
        tile_mac(W, Qc, Y);
        win_resolve_low(W, T);  // all zero by construction
      } else {
        st_block(QR, c - NB, NT, T);
      }
      win_shift(W);
    }
    // value was < R + N: bring it back below R
    if ((W.E[0] | W.O[0] | W.cE[0] | W.cO[0]) != 0) sub_mod(QRp, nmodp, tid);
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
#pragma unroll 1
    for (int v = V - 1; v >= 0; v--) {
      uint4 a = x[v * NT + tid];
      uint4 m = nmod[v];
      if (a.w != m.w) return a.w > m.w;
      if (a.z != m.z) return a.z > m.z;
      if (a.y != m.y) return a.y > m.y;
      if (a.x != m.x) return a.x > m.x;
    }
    return true;  // equal counts as >=
  }
};

}  // namespace ipclb200
