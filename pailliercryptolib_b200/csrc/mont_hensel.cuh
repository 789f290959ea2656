// mont_hensel.cuh -- two-digit (Hensel) Montgomery arithmetic mod p^2 for the
// CRT decrypt (PrivateKey::decryptCRT, ipcl/pri_key.cpp:114-157: the two
// ippModExp calls ct^(p-1) mod p^2, ct^(q-1) mod q^2 and the L function).
//
// A residue X mod p^2 is held as a pair of HALF-WIDTH digits (x0, w),
// 0 <= x0, w < R = 2^(32*LH), LH = words of p:
//
//        X~ = x0 - w*p  (mod p^2),      X~ = X * R mod p^2  (Montgomery form)
//
// The p^2 term of a product vanishes, so a multiply is two half-width CIOS
// sweeps mod p:
//   pass A:  x0*y0 + m*p = TA*R   (quotient digits q_i of m kept)   z0 = TA - ovA*p
//   pass B:  U = x0*wy + wx*y0 + m ;  U + m'*p = V*R ;              wz = V - ovA
// = 5 half-width limb products for a multiply and 4 for a squaring (2*x0*w is
// ONE product with the doubled multiplier) against 8 for the full-width
// Montgomery product mod p^2 the generic kernel performs.  All additions, no
// subtraction: the negative sign of digit 1 absorbs the -m*p of pass A.
// L(x) = (x-1)/p needs no division: for x = c^(p-1) = 1 + l*p the pair is
// (R - p, w) with l = -(w+1)/R mod p, so  L(x)*hp = MontMul_p(w+1, -hp).
// tools/model_hensel.py checks the algebra and the bounds against pow(),
// tools/model_hensel_words.py is the word-level model of the rows below.
//
// Layout: one (ciphertext, side) task = a group of T lanes, K limbs of each
// digit per lane (K*T = LH).  The multiplier limbs of every sweep are read from
// shared memory (4 rows per 128-bit load): a squaring stages x0 and 2w there, a
// window multiply finds its table entry there, prefetched with one TMA bulk copy
// per group (cp.async.bulk + mbarrier) while the squarings before it run.
#pragma once
#include <cstdint>

#include "mont_core.cuh"

namespace ipclb200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// order generic-proxy accesses (st.global of the table, ld.shared of the
// staging area) before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_async_proxy() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// ROWS: CIOS rows per iteration of the (rolled) sweep loops, a multiple of 4.
// ptxas needs ~30 register moves per iteration to bring the two accumulator
// arrays back to their registers, so fewer iterations are cheaper -- until the
// hot loop outgrows the instruction cache.
template <int K, int T, int ROWS = 8, bool W64 = false, bool COMPACT = false>
struct HMont {
  using M = Mont<K, T>;
  static constexpr int LH = K * T;
  static_assert(ROWS % 4 == 0 && LH % ROWS == 0, "ROWS must divide LH");
  // per-group shared-memory area (words).  COMPACT: no separate buffer for the
  // table entry of the next multiply -- it is fetched into S0|S1 when the multiply
  // starts (a squaring's staging areas, dead by then) and digit 0 of a result is
  // parked in its own area P: 4*LH + 4 words per task instead of 5*LH + 4, which
  // is what lets the thread-per-task layout keep 12 warps on an SM.
  static constexpr int kS0 = 0;         // pass A multiplier of a squaring (x0)
  static constexpr int kS1 = LH;        // pass B multiplier of a squaring (2w mod R)
  static constexpr int kSQ = 2 * LH;    // quotient digits of pass A
  static constexpr int kT0 = COMPACT ? 0 : 3 * LH;  // table entry: y0 | wy
  static constexpr int kP = COMPACT ? 3 * LH : 0;   // digit 0 parked during pass B
  // 16-byte multiple, 4-word bank skew per group
  static constexpr int kStride = (COMPACT ? 4 : 5) * LH + 4;

  __device__ __forceinline__ static int lane_t() { return M::lane_t(); }

  // ---- one CIOS row ---------------------------------------------------------
  // MODE 0: pass A          P,Q += a*b ; reduce            (q returned in qout)
  // MODE 1: pass B, square  P,Q += a*b + mi ; reduce
  // MODE 2: pass B, multiply P,Q += a*b + a2*b2 + mi ; reduce
  // mi (lane 0 of the group only, 0 elsewhere) is the quotient digit of pass A
  // for this row.  It is never added physically: it enters the quotient,
  // q = (P0 + mi)*n0', and the limb it completes is P0 + lo(n0*q) + mi = 2^32 *
  // [mi != 0], i.e. a carry of [mi != 0] into limb 1, preset into the odd n*q
  // chain.  P[0] itself is shifted out (lane 0 sends nothing down).
  template <int MODE>
  __device__ __forceinline__ static uint32_t row(
      uint32_t (&P)[K + 1], uint32_t (&Q)[K + 1], const uint32_t (&a)[K],
      const uint32_t (&a2)[K], const uint32_t (&n)[K], uint32_t b, uint32_t b2,
      uint32_t mi, uint32_t in_limb, uint32_t n0inv, uint32_t& qout) {
    uint32_t t0, t1;
    add_cc(t0, Q[K], in_limb);
    addc(t1, 0, 0);
    add_cc(P[0], P[0], Q[1]);
#pragma unroll
    for (int u = 0; u < K / 2 - 1; u++) {
      madc_lo_cc(Q[2 * u], a[2 * u + 1], b, Q[2 * u + 2]);
      madc_hi_cc(Q[2 * u + 1], a[2 * u + 1], b, Q[2 * u + 3]);
    }
    madc_lo_cc(Q[K - 2], a[K - 1], b, t0);
    madc_hi_cc(Q[K - 1], a[K - 1], b, t1);
    addc(Q[K], 0, 0);
    mad_lo_cc(P[0], a[0], b, P[0]);
    madc_hi_cc(P[1], a[0], b, P[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(P[2 * u], a[2 * u], b, P[2 * u]);
      madc_hi_cc(P[2 * u + 1], a[2 * u], b, P[2 * u + 1]);
    }
    addc(P[K], P[K], 0);
    if (MODE == 2) {
      mad_lo_cc(P[0], a2[0], b2, P[0]);
      madc_hi_cc(P[1], a2[0], b2, P[1]);
#pragma unroll
      for (int u = 1; u < K / 2; u++) {
        madc_lo_cc(P[2 * u], a2[2 * u], b2, P[2 * u]);
        madc_hi_cc(P[2 * u + 1], a2[2 * u], b2, P[2 * u + 1]);
      }
      addc(P[K], P[K], 0);
      mad_lo_cc(Q[0], a2[1], b2, Q[0]);
      madc_hi_cc(Q[1], a2[1], b2, Q[1]);
#pragma unroll
      for (int u = 1; u < K / 2; u++) {
        madc_lo_cc(Q[2 * u], a2[2 * u + 1], b2, Q[2 * u]);
        madc_hi_cc(Q[2 * u + 1], a2[2 * u + 1], b2, Q[2 * u + 1]);
      }
      addc(Q[K], Q[K], 0);
    }
    uint32_t q = (MODE == 0 ? P[0] : P[0] + mi) * n0inv;
    if (T > 1) q = __shfl_sync(IPCLB200_FULL_MASK, q, 0, T);
    qout = q;
    mad_lo_cc(P[0], n[0], q, P[0]);
    madc_hi_cc(P[1], n[0], q, P[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(P[2 * u], n[2 * u], q, P[2 * u]);
      madc_hi_cc(P[2 * u + 1], n[2 * u], q, P[2 * u + 1]);
    }
    addc(P[K], P[K], 0);
    if (MODE == 0) {
      mad_lo_cc(Q[0], n[1], q, Q[0]);
    } else {
      uint32_t scratch;
      add_cc(scratch, mi ? 1u : 0u, 0xffffffffu);  // CF = [mi != 0]
      madc_lo_cc(Q[0], n[1], q, Q[0]);
    }
    madc_hi_cc(Q[1], n[1], q, Q[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(Q[2 * u], n[2 * u + 1], q, Q[2 * u]);
      madc_hi_cc(Q[2 * u + 1], n[2 * u + 1], q, Q[2 * u + 1]);
    }
    addc(Q[K], Q[K], 0);
    if (T == 1) return 0u;  // one lane holds the whole digit: nothing crosses
    uint32_t down = __shfl_down_sync(IPCLB200_FULL_MASK, P[0], 1, T);
    return (lane_t() == T - 1) ? 0u : down;
  }

  // ---- the same row on 64-bit accumulator words (W64): array A of K limbs + 1
  // overflow limb = K/2 words w[u] = {A[2u], A[2u+1]} and top = A[K].  In the
  // odd role the array moves down one word per row (w[u] <- w[u+1]), hi(w[0]) is
  // the stranded half that belongs to limb 0.
  template <int MODE>
  __device__ __forceinline__ static uint32_t row64(
      uint64_t (&P)[K / 2], uint32_t& Pt, uint64_t (&Q)[K / 2], uint32_t& Qt,
      const uint32_t (&a)[K], const uint32_t (&a2)[K], const uint32_t (&n)[K], uint32_t b,
      uint32_t b2, uint32_t mi, uint32_t in_limb, uint32_t n0inv, uint32_t& qout) {
    uint64_t t;
    add_wide(t, Qt, in_limb);
    add_lo_cc(P[0], P[0], hi32(Q[0]));
#pragma unroll
    for (int u = 0; u < K / 2 - 1; u++) madc_wide_cc(Q[u], a[2 * u + 1], b, Q[u + 1]);
    madc_wide_cc(Q[K / 2 - 1], a[K - 1], b, t);
    addc(Qt, 0, 0);
    mad_wide_cc(P[0], a[0], b, P[0]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) madc_wide_cc(P[u], a[2 * u], b, P[u]);
    addc(Pt, Pt, 0);
    if (MODE == 2) {
      mad_wide_cc(P[0], a2[0], b2, P[0]);
#pragma unroll
      for (int u = 1; u < K / 2; u++) madc_wide_cc(P[u], a2[2 * u], b2, P[u]);
      addc(Pt, Pt, 0);
      mad_wide_cc(Q[0], a2[1], b2, Q[0]);
#pragma unroll
      for (int u = 1; u < K / 2; u++) madc_wide_cc(Q[u], a2[2 * u + 1], b2, Q[u]);
      addc(Qt, Qt, 0);
    }
    uint32_t q = (MODE == 0 ? lo32(P[0]) : lo32(P[0]) + mi) * n0inv;
    if (T > 1) q = __shfl_sync(IPCLB200_FULL_MASK, q, 0, T);
    qout = q;
    mad_wide_cc(P[0], n[0], q, P[0]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) madc_wide_cc(P[u], n[2 * u], q, P[u]);
    addc(Pt, Pt, 0);
    if (MODE == 0) {
      mad_wide_cc(Q[0], n[1], q, Q[0]);
    } else {
      uint32_t scratch;
      add_cc(scratch, mi ? 1u : 0u, 0xffffffffu);  // CF = [mi != 0]
      madc_wide_cc(Q[0], n[1], q, Q[0]);
    }
#pragma unroll
    for (int u = 1; u < K / 2; u++) madc_wide_cc(Q[u], n[2 * u + 1], q, Q[u]);
    addc(Qt, Qt, 0);
    if (T == 1) return 0u;
    uint32_t down = __shfl_down_sync(IPCLB200_FULL_MASK, lo32(P[0]), 1, T);
    return (lane_t() == T - 1) ? 0u : down;
  }
  __device__ __forceinline__ static void unpack64(uint32_t (&A)[K + 1], const uint64_t (&w)[K / 2],
                                                  uint32_t top) {
#pragma unroll
    for (int u = 0; u < K / 2; u++) {
      A[2 * u] = lo32(w[u]);
      A[2 * u + 1] = hi32(w[u]);
    }
    A[K] = top;
  }

  // E (even role) + O (odd role, one-word shift pending) + the top word (t0,t1)
  // -> K limbs per lane with the cross-lane carries resolved.  Returns how many
  // times R the value overflows (group-uniform).
  __device__ __forceinline__ static uint32_t assemble(uint32_t (&r)[K],
                                                      const uint32_t (&E)[K + 1],
                                                      const uint32_t (&O)[K + 1],
                                                      uint32_t t0, uint32_t t1) {
    uint32_t ov, g;
    add_cc(r[0], E[0], O[1]);
#pragma unroll
    for (int j = 1; j < K - 1; j++) addc_cc(r[j], E[j], O[j + 1]);
    addc_cc(r[K - 1], E[K - 1], t0);
    addc(ov, E[K], t1);
    if (T == 1) return ov;  // the lane's own overflow limb is the multiple of R
    uint32_t ov_in = __shfl_up_sync(IPCLB200_FULL_MASK, ov, 1, T);
    if (lane_t() == 0) ov_in = 0;
    add_cc(r[0], r[0], ov_in);
#pragma unroll
    for (int j = 1; j < K; j++) addc_cc(r[j], r[j], 0);
    addc(g, 0, 0);
    const uint32_t top = M::resolve(r, g);
    return __shfl_sync(IPCLB200_FULL_MASK, ov, T - 1, T) + top;
  }

  // value = r + c*R - dec  ->  r in [0, R) by subtracting multiples of n.
  // c, dec group-uniform.  Returns the number of subtractions (group-uniform).
  // The first round takes the decrement along: r + ~n + (1 - dec), or r + ~0 if
  // there is nothing to subtract; each participating round leaves c - 1 + carry.
  // ONE copy of the add chain (the hot loop has to stay in the instruction cache).
  __device__ __forceinline__ static uint32_t reduce(uint32_t (&r)[K], uint32_t c,
                                                    uint32_t dec,
                                                    const uint32_t (&n)[K]) {
    uint32_t subs = 0;
#pragma unroll 1
    while (__any_sync(IPCLB200_FULL_MASK, (c | dec) != 0)) {
      const uint32_t mc = c ? 0xffffffffu : 0u;           // subtract n
      const uint32_t md = (!c && dec) ? 0xffffffffu : 0u;  // only the decrement
      uint32_t y[K];
#pragma unroll
      for (int j = 0; j < K; j++) y[j] = (~n[j] & mc) | md;
      const uint32_t carry = M::group_add(r, y, (c && !dec) ? 1u : 0u);
      subs += c ? 1u : 0u;
      c = (c | dec) ? c - 1u + carry : c;
      dec = 0;
    }
    return subs;
  }

  // ---- pass A: r = a * B / R mod n (r < R), B = LH words at bs (shared) -------
  // Quotient digits go to qs (shared).  Returns ovA: 1 if n was subtracted.
  __device__ __forceinline__ static uint32_t pass_a(uint32_t (&r)[K],
                                                    const uint32_t (&a)[K],
                                                    const uint32_t* bs, uint32_t* qs,
                                                    const uint32_t (&n)[K],
                                                    uint32_t n0inv) {
    uint32_t E[K + 1], O[K + 1], t0, t1;
    sweep_a(E, O, t0, t1, a, bs, qs, n, n0inv);
    const uint32_t c = assemble(r, E, O, t0, t1);
    reduce(r, c, 0u, n);
    return c;
  }
  // the rows of pass A: leaves the accumulator arrays and the top word (t0,t1)
  __device__ __forceinline__ static void sweep_a(uint32_t (&E)[K + 1], uint32_t (&O)[K + 1],
                                                 uint32_t& t0, uint32_t& t1,
                                                 const uint32_t (&a)[K],
                                                 const uint32_t* bs, uint32_t* qs,
                                                 const uint32_t (&n)[K],
                                                 uint32_t n0inv) {
    uint32_t in_limb = 0;
    const bool l0 = lane_t() == 0;
    if (W64) {
      uint64_t Ew[K / 2], Ow[K / 2];
      uint32_t Et = 0, Ot = 0;
#pragma unroll
      for (int u = 0; u < K / 2; u++) {
        Ew[u] = 0;
        Ow[u] = 0;
      }
#pragma unroll 1
      for (int i = 0; i < LH; i += ROWS) {
#pragma unroll
        for (int k = 0; k < ROWS; k += 4) {
          const uint4 bv = *reinterpret_cast<const uint4*>(bs + i + k);
          uint4 qv;
          in_limb = row64<0>(Ew, Et, Ow, Ot, a, a, n, bv.x, 0u, 0u, in_limb, n0inv, qv.x);
          in_limb = row64<0>(Ow, Ot, Ew, Et, a, a, n, bv.y, 0u, 0u, in_limb, n0inv, qv.y);
          in_limb = row64<0>(Ew, Et, Ow, Ot, a, a, n, bv.z, 0u, 0u, in_limb, n0inv, qv.z);
          in_limb = row64<0>(Ow, Ot, Ew, Et, a, a, n, bv.w, 0u, 0u, in_limb, n0inv, qv.w);
          if (l0) *reinterpret_cast<uint4*>(qs + i + k) = qv;
        }
      }
      unpack64(E, Ew, Et);
      unpack64(O, Ow, Ot);
    } else {
#pragma unroll
      for (int j = 0; j <= K; j++) {
        E[j] = 0;
        O[j] = 0;
      }
#pragma unroll 1
      for (int i = 0; i < LH; i += ROWS) {
#pragma unroll
        for (int k = 0; k < ROWS; k += 4) {
          const uint4 bv = *reinterpret_cast<const uint4*>(bs + i + k);
          uint4 qv;
          in_limb = row<0>(E, O, a, a, n, bv.x, 0u, 0u, in_limb, n0inv, qv.x);
          in_limb = row<0>(O, E, a, a, n, bv.y, 0u, 0u, in_limb, n0inv, qv.y);
          in_limb = row<0>(E, O, a, a, n, bv.z, 0u, 0u, in_limb, n0inv, qv.z);
          in_limb = row<0>(O, E, a, a, n, bv.w, 0u, 0u, in_limb, n0inv, qv.w);
          if (l0) *reinterpret_cast<uint4*>(qs + i + k) = qv;
        }
      }
    }
    add_cc(t0, O[K], in_limb);
    addc(t1, 0, 0);
  }

  // ---- pass B: r = (a*B1 [+ a2*B0] + m) / R [+ hb*a] - dec mod n (r < R) -------
  // B1, B0 (multiplier streams) and m = qs are shared-memory words.  hb: the
  // bit 2w lost when it was truncated to LH words (squaring only): + a*R before
  // the division, added as one unreduced half row at the end.
  template <bool TWO>
  __device__ __forceinline__ static void pass_b(
      uint32_t (&r)[K], const uint32_t (&a)[K], const uint32_t (&a2)[K],
      const uint32_t* b1s, const uint32_t* b0s, const uint32_t* qs,
      const uint32_t (&n)[K], uint32_t n0inv, uint32_t hb, uint32_t dec) {
    uint32_t E[K + 1], O[K + 1], t0, t1;
    sweep_b<TWO>(E, O, t0, t1, a, a2, b1s, b0s, qs, n, n0inv, hb);
    const uint32_t c = assemble(r, E, O, t0, t1);
    reduce(r, c, dec, n);
  }
  template <bool TWO>
  __device__ __forceinline__ static void sweep_b(
      uint32_t (&E)[K + 1], uint32_t (&O)[K + 1], uint32_t& t0, uint32_t& t1,
      const uint32_t (&a)[K], const uint32_t (&a2)[K], const uint32_t* b1s,
      const uint32_t* b0s, const uint32_t* qs, const uint32_t (&n)[K], uint32_t n0inv,
      uint32_t hb) {
    uint32_t in_limb = 0, qd;
    const bool l0 = lane_t() == 0;
    constexpr int MD = TWO ? 2 : 1;
    __syncwarp();  // qs was written by lane 0 of the group
    if (W64) {
      uint64_t Ew[K / 2], Ow[K / 2];
      uint32_t Et = 0, Ot = 0;
#pragma unroll
      for (int u = 0; u < K / 2; u++) {
        Ew[u] = 0;
        Ow[u] = 0;
      }
#pragma unroll 1
      for (int i = 0; i < LH; i += ROWS) {
#pragma unroll
        for (int k = 0; k < ROWS; k += 4) {
          const uint4 bv = *reinterpret_cast<const uint4*>(b1s + i + k);
          uint4 cv = make_uint4(0, 0, 0, 0);
          if (TWO) cv = *reinterpret_cast<const uint4*>(b0s + i + k);
          // only the group's lane 0 injects the quotient digits (it also wrote them:
          // program order, no other lane ever touches qs)
          uint4 mv = make_uint4(0, 0, 0, 0);
          if (l0) mv = *reinterpret_cast<const uint4*>(qs + i + k);
          in_limb = row64<MD>(Ew, Et, Ow, Ot, a, a2, n, bv.x, cv.x, mv.x, in_limb, n0inv, qd);
          in_limb = row64<MD>(Ow, Ot, Ew, Et, a, a2, n, bv.y, cv.y, mv.y, in_limb, n0inv, qd);
          in_limb = row64<MD>(Ew, Et, Ow, Ot, a, a2, n, bv.z, cv.z, mv.z, in_limb, n0inv, qd);
          in_limb = row64<MD>(Ow, Ot, Ew, Et, a, a2, n, bv.w, cv.w, mv.w, in_limb, n0inv, qd);
        }
      }
      unpack64(E, Ew, Et);
      unpack64(O, Ow, Ot);
    } else {
#pragma unroll
      for (int j = 0; j <= K; j++) {
        E[j] = 0;
        O[j] = 0;
      }
#pragma unroll 1
      for (int i = 0; i < LH; i += ROWS) {
#pragma unroll
        for (int k = 0; k < ROWS; k += 4) {
          const uint4 bv = *reinterpret_cast<const uint4*>(b1s + i + k);
          uint4 cv = make_uint4(0, 0, 0, 0);
          if (TWO) cv = *reinterpret_cast<const uint4*>(b0s + i + k);
          // only the group's lane 0 injects the quotient digits (it also wrote them:
          // program order, no other lane ever touches qs)
          uint4 mv = make_uint4(0, 0, 0, 0);
          if (l0) mv = *reinterpret_cast<const uint4*>(qs + i + k);
          in_limb = row<MD>(E, O, a, a2, n, bv.x, cv.x, mv.x, in_limb, n0inv, qd);
          in_limb = row<MD>(O, E, a, a2, n, bv.y, cv.y, mv.y, in_limb, n0inv, qd);
          in_limb = row<MD>(E, O, a, a2, n, bv.z, cv.z, mv.z, in_limb, n0inv, qd);
          in_limb = row<MD>(O, E, a, a2, n, bv.w, cv.w, mv.w, in_limb, n0inv, qd);
        }
      }
    }
    add_cc(t0, O[K], in_limb);
    addc(t1, 0, 0);
    if (!TWO) half_row(E, O, t0, t1, a, hb);
  }
  // half row: E,O += a * hb (no quotient, no shift).  Odd word u of O lives in
  // O[2u+2..2u+3], the top one in (t0,t1).
  __device__ __forceinline__ static void half_row(uint32_t (&E)[K + 1], uint32_t (&O)[K + 1],
                                                  uint32_t& t0, uint32_t& t1,
                                                  const uint32_t (&a)[K], uint32_t hb) {
    mad_lo_cc(O[2], a[1], hb, O[2]);
    madc_hi_cc(O[3], a[1], hb, O[3]);
#pragma unroll
    for (int u = 1; u < K / 2 - 1; u++) {
      madc_lo_cc(O[2 * u + 2], a[2 * u + 1], hb, O[2 * u + 2]);
      madc_hi_cc(O[2 * u + 3], a[2 * u + 1], hb, O[2 * u + 3]);
    }
    madc_lo_cc(t0, a[K - 1], hb, t0);
    madc_hi_cc(t1, a[K - 1], hb, t1);
    mad_lo_cc(E[0], a[0], hb, E[0]);
    madc_hi_cc(E[1], a[0], hb, E[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(E[2 * u], a[2 * u], hb, E[2 * u]);
      madc_hi_cc(E[2 * u + 1], a[2 * u], hb, E[2 * u + 1]);
    }
    addc(E[K], E[K], 0);
  }

  // this lane's K limbs -> the group's LH-word shared operand
  __device__ __forceinline__ static void put(uint32_t* dst, const uint32_t (&x)[K]) {
    uint32_t* d = dst + lane_t() * K;
    if (K % 4 == 0) {
#pragma unroll
      for (int j = 0; j < K; j += 4)
        *reinterpret_cast<uint4*>(d + j) = make_uint4(x[j], x[j + 1], x[j + 2], x[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < K; j += 2)
        *reinterpret_cast<uint2*>(d + j) = make_uint2(x[j], x[j + 1]);
    }
  }

  // ---- (x0, w) <- (x0, w)^2 -------------------------------------------------
  __device__ __forceinline__ static void sqr(uint32_t (&x0)[K], uint32_t (&w)[K],
                                             uint32_t* sm, const uint32_t (&n)[K],
                                             uint32_t n0inv) {
    uint32_t hb;
    {
      uint32_t dw[K];
      uint32_t below = 0;
      if (T > 1) {
        below = __shfl_up_sync(IPCLB200_FULL_MASK, w[K - 1], 1, T);
        if (lane_t() == 0) below = 0;
      }
      dw[0] = __funnelshift_l(below, w[0], 1);
#pragma unroll
      for (int j = 1; j < K; j++) dw[j] = __funnelshift_l(w[j - 1], w[j], 1);
      hb = w[K - 1] >> 31;
      if (T > 1) hb = __shfl_sync(IPCLB200_FULL_MASK, hb, T - 1, T);
      __syncwarp();
      put(sm + kS0, x0);
      put(sm + kS1, dw);
      __syncwarp();
    }
    uint32_t z0[K];
    const uint32_t ovA = pass_a(z0, x0, sm + kS0, sm + kSQ, n, n0inv);
    pass_b<false>(w, x0, x0, sm + kS1, sm + kS1, sm + kSQ, n, n0inv, hb, ovA);
#pragma unroll
    for (int j = 0; j < K; j++) x0[j] = z0[j];
  }

  // ---- (x0, w) <- (x0, w) * (y0, wy), the pair y at ys (shared: y0 | wy) -------
  __device__ __forceinline__ static void mul(uint32_t (&x0)[K], uint32_t (&w)[K],
                                             const uint32_t* ys, uint32_t* sm,
                                             const uint32_t (&n)[K], uint32_t n0inv) {
    uint32_t z0[K], wz[K];
    const uint32_t ovA = pass_a(z0, x0, ys, sm + kSQ, n, n0inv);
    pass_b<true>(wz, x0, w, ys + LH, ys, sm + kSQ, n, n0inv, 0u, ovA);
#pragma unroll
    for (int j = 0; j < K; j++) {
      x0[j] = z0[j];
      w[j] = wz[j];
    }
  }

  // ---- one step of the schedule through ONE copy of pass A: a squaring, or a
  //      multiply by the pair at ys (the hot loop stays in the instruction cache)
  __device__ __forceinline__ static void step(uint32_t (&x0)[K], uint32_t (&w)[K],
                                              bool is_mul, const uint32_t* ys,
                                              uint32_t* sm, const uint32_t (&n)[K],
                                              uint32_t n0inv) {
    uint32_t hb = 0;
    if (!is_mul) {
      uint32_t dw[K];
      uint32_t below = 0;
      if (T > 1) {
        below = __shfl_up_sync(IPCLB200_FULL_MASK, w[K - 1], 1, T);
        if (lane_t() == 0) below = 0;
      }
      dw[0] = __funnelshift_l(below, w[0], 1);
#pragma unroll
      for (int j = 1; j < K; j++) dw[j] = __funnelshift_l(w[j - 1], w[j], 1);
      hb = w[K - 1] >> 31;
      if (T > 1) hb = __shfl_sync(IPCLB200_FULL_MASK, hb, T - 1, T);
      __syncwarp();
      put(sm + kS0, x0);
      put(sm + kS1, dw);
      __syncwarp();
    }
    uint32_t ovA;
    {
      // digit 0 of the result waits in this lane's slice of S0 (dead after pass A:
      // a squaring has consumed its copy of x0, a multiply reads the entry buffer)
      // while pass B runs: K registers less in the hottest loop
      uint32_t z0[K];
      ovA = pass_a(z0, x0, is_mul ? ys : sm + kS0, sm + kSQ, n, n0inv);
      __syncwarp();  // every lane of the group is done reading S0
      put(sm + kP, z0);
    }
    if (is_mul) {
      uint32_t wz[K];
      pass_b<true>(wz, x0, w, ys + LH, ys, sm + kSQ, n, n0inv, 0u, ovA);
#pragma unroll
      for (int j = 0; j < K; j++) w[j] = wz[j];
    } else {
      pass_b<false>(w, x0, x0, sm + kS1, sm + kS1, sm + kSQ, n, n0inv, hb, ovA);
    }
    M::load(x0, sm + kP);  // this lane's own slice: no synchronisation needed
  }

  // ---- (x0, w) <- (a, 0) * (k0, kw), the constant pair already staged at
  //      sm+kS0 | sm+kS1 (prologue) ---------------------------------------------
  __device__ __forceinline__ static void mul_digit(uint32_t (&z0)[K], uint32_t (&wz)[K],
                                                   const uint32_t (&a)[K], uint32_t* sm,
                                                   const uint32_t (&n)[K],
                                                   uint32_t n0inv) {
    const uint32_t ovA = pass_a(z0, a, sm + kS0, sm + kSQ, n, n0inv);
    pass_b<false>(wz, a, a, sm + kS1, sm + kS1, sm + kSQ, n, n0inv, 0u, ovA);
  }

  // ---- one row of a PLAIN product (no reduction): P,Q += a*b, the lowest limb
  //      leaves the accumulator (lane 0: a finished limb of the product) ---------
  __device__ __forceinline__ static uint32_t row_plain(uint32_t (&P)[K + 1],
                                                       uint32_t (&Q)[K + 1],
                                                       const uint32_t (&a)[K], uint32_t b,
                                                       uint32_t in_limb, uint32_t& out) {
    uint32_t t0, t1;
    add_cc(t0, Q[K], in_limb);
    addc(t1, 0, 0);
    add_cc(P[0], P[0], Q[1]);
#pragma unroll
    for (int u = 0; u < K / 2 - 1; u++) {
      madc_lo_cc(Q[2 * u], a[2 * u + 1], b, Q[2 * u + 2]);
      madc_hi_cc(Q[2 * u + 1], a[2 * u + 1], b, Q[2 * u + 3]);
    }
    madc_lo_cc(Q[K - 2], a[K - 1], b, t0);
    madc_hi_cc(Q[K - 1], a[K - 1], b, t1);
    addc(Q[K], 0, 0);
    mad_lo_cc(P[0], a[0], b, P[0]);
    madc_hi_cc(P[1], a[0], b, P[1]);
#pragma unroll
    for (int u = 1; u < K / 2; u++) {
      madc_lo_cc(P[2 * u], a[2 * u], b, P[2 * u]);
      madc_hi_cc(P[2 * u + 1], a[2 * u], b, P[2 * u + 1]);
    }
    addc(P[K], P[K], 0);
    out = P[0];
    if (T == 1) return 0u;
    uint32_t down = __shfl_down_sync(IPCLB200_FULL_MASK, P[0], 1, T);
    return (lane_t() == T - 1) ? 0u : down;
  }

  // hi:lo = a * B (2*LH words), B = LH words at bs (shared).  The low LH words
  // are written to los (shared, by lane 0 of the group), the high ones returned
  // in hi (K limbs per lane).
  __device__ __forceinline__ static void mul_plain(uint32_t (&hi)[K], const uint32_t (&a)[K],
                                                   const uint32_t* bs, uint32_t* los) {
    uint32_t E[K + 1], O[K + 1];
#pragma unroll
    for (int j = 0; j <= K; j++) {
      E[j] = 0;
      O[j] = 0;
    }
    uint32_t in_limb = 0;
    const bool l0 = lane_t() == 0;
#pragma unroll 1
    for (int i = 0; i < LH; i += 4) {
      const uint4 bv = *reinterpret_cast<const uint4*>(bs + i);
      uint4 ov;
      in_limb = row_plain(E, O, a, bv.x, in_limb, ov.x);
      in_limb = row_plain(O, E, a, bv.y, in_limb, ov.y);
      in_limb = row_plain(E, O, a, bv.z, in_limb, ov.z);
      in_limb = row_plain(O, E, a, bv.w, in_limb, ov.w);
      if (l0) *reinterpret_cast<uint4*>(los + i) = ov;
    }
    uint32_t t0, t1;
    add_cc(t0, O[K], in_limb);
    addc(t1, 0, 0);
    assemble(hi, E, O, t0, t1);  // a*B < R^2: nothing above the high half
  }

  // ---- (x0, w) = sum_j (c_j, 0) * K_j : c_j = the j-th LH-word chunk at c (global),
  //      K_j = the pair at consts + 2*j*LH (global).  With K_j = R^(j+1) in
  //      Montgomery form this is c * R mod p^2 for the integer c = sum c_j R^j ------
  __device__ __forceinline__ static void enter(uint32_t (&x0)[K], uint32_t (&w)[K],
                                               const uint32_t* c, int nchunks,
                                               const uint32_t* consts, uint32_t* sm,
                                               const uint32_t (&n)[K], uint32_t n0inv) {
#pragma unroll 1
    for (int j = 0; j < nchunks; j++) {
      uint32_t a[K], z0[K], wz[K];
      __syncwarp();
      M::load(a, consts + (size_t)(2 * j) * LH);
      put(sm + kS0, a);
      M::load(a, consts + (size_t)(2 * j + 1) * LH);
      put(sm + kS1, a);
      __syncwarp();
      M::load(a, c + (size_t)j * LH);
      mul_digit(z0, wz, a, sm, n, n0inv);
      if (j == 0) {
#pragma unroll
        for (int k = 0; k < K; k++) {
          x0[k] = z0[k];
          w[k] = wz[k];
        }
      } else {
        add(x0, w, z0, wz, n);
      }
    }
  }

  // ---- plain pair -> canonical residue: for u = u0 - uw*n (mod n^2) and an
  //      extra digit-1 term `am` (< n, canonical): with a = u0 mod n (u0 = a + c*n)
  //          u + am*n = a + ((am + c - uw) mod n) * n        (mod n^2)
  //      lo | hi = the 2*LH-word result.  S0 and SQ of the group area are scratch.
  __device__ __forceinline__ static void to_canonical(uint32_t (&lo)[K], uint32_t (&hi)[K],
                                                      uint32_t (&u0)[K], uint32_t (&uw)[K],
                                                      uint32_t (&am)[K], uint32_t* sm,
                                                      const uint32_t (&n)[K]) {
    const uint32_t c = M::sub_n_if_ge(u0, n);
    M::sub_n_if_ge(uw, n);
    {
      uint32_t y[K];
#pragma unroll
      for (int j = 0; j < K; j++) y[j] = 0;
      M::group_add(am, y, c);  // am + c <= n
      M::sub_n_if_ge(am, n);
#pragma unroll
      for (int j = 0; j < K; j++) y[j] = ~uw[j];
      const uint32_t nb = M::group_add(am, y, 1u);  // - uw; carry out <=> no borrow
#pragma unroll
      for (int j = 0; j < K; j++) y[j] = nb ? 0u : n[j];
      M::group_add(am, y, 0u);  // + n after a borrow (the carry out cancels it)
    }
    __syncwarp();
    put(sm + kS0, am);
    __syncwarp();
    mul_plain(hi, n, sm + kS0, sm + kSQ);
    __syncwarp();
    M::load(lo, sm + kSQ);  // this lane's K low limbs
    const uint32_t cy = M::group_add(lo, u0, 0u);
    uint32_t y[K];
#pragma unroll
    for (int j = 0; j < K; j++) y[j] = 0;
    M::group_add(hi, y, cy);
  }

  // ---- (x0, w) += (y0, wy)  (prologue only) -----------------------------------
  // Each p taken off digit 0 is one taken off w, i.e. + (p - 1).
  __device__ __forceinline__ static void add(uint32_t (&x0)[K], uint32_t (&w)[K],
                                             const uint32_t (&y0)[K],
                                             const uint32_t (&wy)[K],
                                             const uint32_t (&n)[K]) {
    const uint32_t c0 = M::group_add(x0, y0, 0u);
    const uint32_t subs = reduce(x0, c0, 0u, n);
    uint32_t c = M::group_add(w, wy, 0u);
#pragma unroll 1
    for (uint32_t it = 0; it < 3; it++) {
      if (!__any_sync(IPCLB200_FULL_MASK, subs > it)) break;
      uint32_t y[K];
#pragma unroll
      for (int j = 0; j < K; j++) y[j] = subs > it ? n[j] : 0u;
      if (lane_t() == 0 && subs > it) y[0] -= 1u;  // n is odd: no borrow
      c += M::group_add(w, y, 0u);
    }
    reduce(w, c, 0u, n);
  }
};

}  // namespace ipclb200
