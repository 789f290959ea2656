// kernel_hensel_decrypt.cuh -- K4h, the CRT decrypt in two-digit (Hensel)
// arithmetic.  A header of its own because the kernel is instantiated in its
// own translation unit (hensel_decrypt.cu, see hensel_launch.hpp).
#pragma once
#include <cstdint>

#include "kernels_common.cuh"
#include "mont_hensel.cuh"

namespace ipclb200 {

// --------------------------------------------------------------------------
// K4h: CRT decrypt in two-digit (Hensel) arithmetic, mont_hensel.cuh.  Same job
//     as K4 with half-width digits: per (ciphertext, side) task a group of T
//     lanes, LH = K*T = words of p.  Prologue: the four LH-word chunks c_j of the
//     ciphertext times the key constants R^(j+1) (as pairs) and summed = ct*R mod
//     p^2 as a pair.  Odd-power table (pairs, 2*LH words each) in the group's
//     workspace slot, the sliding-window schedule of p-1 as (squarings, entry)
//     runs; the entry of the next multiply is fetched into shared memory by a TMA
//     bulk copy while the squarings run.  Epilogue: mp = MontMul_p(w+1, -hp),
//     canonical -- the L function and the multiplication by hp in one half-width
//     product.  crt_combine_kernel finishes.
// --------------------------------------------------------------------------
struct HenselSide {
  const uint32_t* blk;   // p | (k0_j | kw_j), j = 0..3 | -hp mod p    (10*LH words)
  // [entries | all_powers << 8, first, (run << 8 | entry)..., (run << 8 | 0xff)]
  const uint32_t* sched;
  uint32_t n0inv;
};

struct DecryptHenselParams {
  const uint32_t* ct;  // count x 4*LH words
  HenselSide s0, s1;
  uint32_t* mpq;  // out: count x 2 x LH words
  size_t count;
  uint32_t* table_ws;  // per group: table_entries x 2*LH words
  int table_entries;
  unsigned int* work_counter;
};

template <int K, int T, bool COMPACT = false>
constexpr size_t hensel_smem_bytes(int threads) {
  return (size_t)(threads / T) * HMont<K, T, 4, false, COMPACT>::kStride * sizeof(uint32_t) +
         (size_t)(threads / 32) * sizeof(uint64_t);
}

template <int K, int T, int ROWS, bool W64, bool COMPACT>
__device__ __forceinline__ void decrypt_hensel_body(const DecryptHenselParams& p);

template <int K, int T, int MINB, int ROWS, bool W64 = false, int BT = kBlockThreads,
          bool COMPACT = false>
__global__ void __launch_bounds__(BT, MINB)
    decrypt_hensel_kernel(const DecryptHenselParams p) {
  decrypt_hensel_body<K, T, ROWS, W64, COMPACT>(p);
}

template <int K, int T, int ROWS, bool W64, bool COMPACT>
__device__ __forceinline__ void decrypt_hensel_body(const DecryptHenselParams& p) {
  using M = Mont<K, T>;
  using H = HMont<K, T, ROWS, W64, COMPACT>;
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
  uint32_t* tab = p.table_ws + gid * ((size_t)2 * LH * p.table_entries);
  constexpr uint32_t kEntryBytes = 2 * LH * sizeof(uint32_t);
  const unsigned int nchunks = (unsigned int)((p.count + GW - 1) / GW);
  for (;;) {
    const unsigned int wk = claim_chunk(p.work_counter);
    if (wk >= 2u * nchunks) break;
    const int side = (int)(wk & 1u);
    const size_t inst = (size_t)(wk >> 1) * GW + (threadIdx.x & 31) / T;
    const bool valid = inst < p.count;
    const size_t ii = valid ? inst : p.count - 1;
    const uint32_t* blk = side ? p.s1.blk : p.s0.blk;
    const uint32_t* sched = side ? p.s1.sched : p.s0.sched;
    const uint32_t n0inv = side ? p.s1.n0inv : p.s0.n0inv;
    uint32_t n[K];
    M::load(n, blk);
    uint32_t x0[K], w[K];
    // ---- prologue: ct -> (x0, w) = ct * R mod p^2 -----------------------------
    H::enter(x0, w, p.ct + ii * (size_t)(4 * LH), 4, blk + LH, sm, n, n0inv);
    // ---- the table in the workspace slot: odd powers x, x^3, ... (sliding
    // window) or, for the constant schedule, all powers x^0 .. x^(n-1) -----------
    const uint32_t tab_word = __ldg(sched);
    const int nodd = (int)(tab_word & 0xffu);
    const bool all_powers = (tab_word >> 8) != 0;
    if (all_powers) {
      // x^0 = 1 in Montgomery form: R mod p^2 = (R - p) + 1*p, i.e. the pair
      // (R - p, p - 1); p is odd, so neither +1 nor -1 carries
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
      M::store(tab + 3 * LH, w);
      __syncwarp();
      H::put(sm + H::kT0, x0);
      H::put(sm + H::kT0 + LH, w);
      __syncwarp();
#pragma unroll 1
      for (int i = 2; i < nodd; i++) {
        H::step(x0, w, true, sm + H::kT0, sm, n, n0inv);  // x^i = x^(i-1) * x
        M::store(tab + (size_t)i * 2 * LH, x0);
        M::store(tab + (size_t)i * 2 * LH + LH, w);
      }
    } else {
      M::store(tab, x0);
      M::store(tab + LH, w);
      if (nodd > 1) {
        // x^2 into the table-entry buffer, then x^(2i+1) = x^(2i-1) * x^2
#pragma unroll 1
        for (int i = 0; i < nodd; i++) {
          H::step(x0, w, i != 0, sm + H::kT0, sm, n, n0inv);
          if (i == 0) {
            __syncwarp();
            H::put(sm + H::kT0, x0);
            H::put(sm + H::kT0 + LH, w);
            __syncwarp();
            M::load(x0, tab);
            M::load(w, tab + LH);
          } else {
            M::store(tab + (size_t)i * 2 * LH, x0);
            M::store(tab + (size_t)i * 2 * LH + LH, w);
          }
        }
      }
    }
    // ---- the schedule -----------------------------------------------------------
    {
      const uint32_t first = __ldg(sched + 1);
      M::load(x0, tab + (size_t)first * 2 * LH);
      M::load(w, tab + (size_t)first * 2 * LH + LH);
    }
    const uint32_t* op = sched + 2;
    uint32_t cur = __ldg(op);
    // table stores (generic proxy) before the TMA reads (async proxy); the last
    // generic reads of the staging area before the TMA overwrites it.  The entry of
    // the next multiply is fetched while the squarings before it run -- or, in the
    // COMPACT layout, into the (by then dead) staging areas of the last squaring
    // when the multiply starts: one exposed L2 round trip per multiply, ~0.3 % of
    // the schedule, for 128 bytes less shared memory per task.
    auto fetch_entry = [&](uint32_t entry) {
      __syncwarp();
      fence_async_proxy();
      if (warp_leader) mbar_expect_tx(bar, GW * kEntryBytes);
      __syncwarp();
      if (group_leader)
        bulk_g2s(sm + H::kT0, tab + (size_t)entry * 2 * LH, kEntryBytes, bar);
    };
    if (!COMPACT && (cur & 0xffu) != 0xffu) fetch_entry(cur & 0xffu);
    uint32_t run = cur >> 8;
#pragma unroll 1
    for (;;) {
      const bool is_mul = run == 0;
      if (is_mul) {
        if ((cur & 0xffu) == 0xffu) break;
        if (COMPACT) fetch_entry(cur & 0xffu);
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
      H::step(x0, w, is_mul, sm + H::kT0, sm, n, n0inv);
      if (is_mul) {
        cur = __ldg(++op);
        run = cur >> 8;
        if (!COMPACT && (cur & 0xffu) != 0xffu) fetch_entry(cur & 0xffu);
      } else {
        run--;
      }
    }
    // ---- epilogue: mp = MontMul_p(w + 1, -hp), canonical ----------------------------
    {
      uint32_t zero[K], t[K];
#pragma unroll
      for (int j = 0; j < K; j++) zero[j] = 0;
      const uint32_t cy = M::group_add(w, zero, 1u);
      if (__any_sync(IPCLB200_FULL_MASK, cy)) M::cond_sub_n(w, n, cy);  // w+1 == R -> R - p
      __syncwarp();
      M::load(t, blk + (size_t)9 * LH);
      H::put(sm + H::kS0, t);
      __syncwarp();
      H::pass_a(t, w, sm + H::kS0, sm + H::kSQ, n, n0inv);
      M::sub_n_if_ge(t, n);
      if (valid) M::store(p.mpq + (inst * 2 + side) * LH, t);
    }
    __syncwarp();
  }
}

}  // namespace ipclb200
